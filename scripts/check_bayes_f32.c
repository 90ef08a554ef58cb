/*
 * check_bayes_f32.c -- exhaustive check of the f32 restatement of  (float)(0.825 * (double)m)  used by synd_bayes_sel
 * (csrc/nbldpc_synd.cuh): over ALL non-negative finite floats m, the formula  fma(m, c1, m * c2)  with c1 = (float)0.825, c2 = (float)(0.825 - c1)
 * must give the reference's bits for every m >= 2^-96 (and for 0); the kernel keeps the double multiplication below that.
 *
 *   gcc -O2 -fopenmp -ffp-contract=off -o /tmp/check_bayes_f32 scripts/check_bayes_f32.c -lm && /tmp/check_bayes_f32
 *   (about a minute on 16 cores; prints the number of mismatches at or above the threshold, which must be 0)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(void)
{
    const double c = 0.825;
    const float c1 = (float)c, c2 = (float)(c - (double)c1);
    const uint32_t threshold = 0x0f800000u;                    /* 2^-96 */
    long bad_above = 0, bad_below = 0;
    printf("c1 = %a  c2 = %a  rest = %a\n", c1, c2, c - (double)c1 - (double)c2);
#pragma omp parallel for reduction(+ : bad_above, bad_below) schedule(static)
    for (long u = 0; u < 0x7f800000L; u++) {
        const float m = u2f((uint32_t)u);
        const float ref = (float)(c * (double)m);
        const float r = fmaf(m, c1, m * c2);
        if (memcmp(&r, &ref, 4)) { if ((uint32_t)u >= threshold || u == 0) bad_above++; else bad_below++; }
    }
    printf("mismatches: %ld at m = 0 or m >= 2^-96 (must be 0), %ld below (handled in f64 by the kernel)\n", bad_above, bad_below);
    return bad_above != 0;
}
