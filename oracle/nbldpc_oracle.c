/*
 * oracle/nbldpc_oracle.c -- TEST INFRASTRUCTURE ONLY (see nbldpc_oracle.h for the rules).
 *
 * CPU restatement of the reference decode path.  Each function cites the reference lines it follows.
 * Arithmetic is kept operation-for-operation (single f32 adds/subs, the f32/f64 mix of channel.c:73,
 * strict '<' scans from the +1e5 sentinel) so that results are bit-identical to the reference;
 * compile with -ffp-contract=off.
 */
#include "nbldpc_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SENT 1e5f

/* ------------------------------------------------------------------------------------------------
 * GF(2^m) tables.  struct.h:117/145/217 give the primitive polynomials; symbol 0 is the zero
 * element and symbol k>=1 is alpha^(k-1), bits stored LSB first (struct.h:119-476).
 * ADD: XOR of binary images then linear search (init.c:37-53); MUL/DIV: init.c:65-130.
 * ---------------------------------------------------------------------------------------------- */
static int build_tables(nbo_code *c)
{
    int GF = c->GF, lg = c->logGF, poly, k, l, i, j;
    if (GF == 16) poly = 0x13; else if (GF == 64) poly = 0x43; else if (GF == 256) poly = 0x11D;
    else return -1;                                  /* init.c:431-435 exits for any other field */
    c->bingf = calloc((size_t)GF * lg, sizeof(int));
    c->addgf = calloc((size_t)GF * GF, sizeof(int));
    c->mulgf = calloc((size_t)GF * GF, sizeof(int));
    c->divgf = calloc((size_t)GF * GF, sizeof(int));
    int *img = calloc(GF, sizeof(int)), *inv = calloc(GF, sizeof(int));
    int x = 1;
    img[0] = 0;
    for (k = 1; k < GF; k++) { img[k] = x; x <<= 1; if (x & GF) x ^= poly; }
    for (k = 0; k < GF; k++) { inv[img[k]] = k; for (l = 0; l < lg; l++) c->bingf[k * lg + l] = (img[k] >> l) & 1; }
    for (i = 0; i < GF; i++) for (j = 0; j < GF; j++) c->addgf[i * GF + j] = inv[img[i] ^ img[j]];
    /* Table_Mul_GF, init.c:65-88 */
    for (i = 0; i < GF; i++) for (j = 0; j < GF; j++) {
        int v;
        if (i == 0 || j == 0) v = 0;
        else if (i == 1) v = j;
        else if (j == 1) v = i;
        else { int t = i + j - 2; v = (t < GF - 1) ? t + 1 : (t % (GF - 1)) + 1; }
        c->mulgf[i * GF + j] = v;
    }
    /* Table_Div_GF, init.c:100-130: a running counter nb that is decremented on the generic entries */
    int nb = GF - 1;
    for (i = 0; i < GF; i++) for (j = 0; j < GF; j++) {
        int v;
        if (j == 0) v = -1;
        else if (i == 0) v = 0;
        else if (j == 1) v = i;
        else v = nb--;
        if (nb < 1) nb = GF - 1;
        c->divgf[i * GF + j] = v;
    }
    free(img); free(inv);
    return 0;
}

/* LoadCode, init.c:143-272 (both alist dialects) */
nbo_code *nbo_load(const char *path, int dialect)
{
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    nbo_code *c = calloc(1, sizeof *c);
    int n, m, k;
    if (fscanf(f, "%d %d %d", &c->N, &c->M, &c->GF) != 3) { fclose(f); free(c); return NULL; }
    c->logGF = (int)rint(log((double)c->GF) / log(2.0));
    c->K = c->N - c->M;
    c->rate = (float)(c->N - c->M) / c->N;
    for (n = 0; n < c->N; n++) { int d; if (fscanf(f, "%d", &d) != 1) goto bad; }
    c->row_deg = calloc(c->M, sizeof(int));
    c->row_ptr = calloc(c->M + 1, sizeof(int));
    for (m = 0; m < c->M; m++) {
        if (fscanf(f, "%d", &c->row_deg[m]) != 1) goto bad;
        c->row_ptr[m + 1] = c->row_ptr[m] + c->row_deg[m];
        if (c->row_deg[m] > c->dc_max) c->dc_max = c->row_deg[m];
    }
    c->E = c->row_ptr[c->M];
    c->col = calloc(c->E, sizeof(int));
    c->val = calloc(c->E, sizeof(int));
    if (dialect == 2) {            /* KN: interleaved (col 1-based, exponent) pairs, init.c:214-224 */
        for (k = 0; k < c->E; k++) {
            int a, b;
            if (fscanf(f, "%d %d", &a, &b) != 2) goto bad;
            c->col[k] = a - 1; c->val[k] = b + 1;
        }
    } else {                        /* UBS: all columns then all coefficients, init.c:197-205 */
        for (k = 0; k < c->E; k++) if (fscanf(f, "%d", &c->col[k]) != 1) goto bad;
        for (k = 0; k < c->E; k++) if (fscanf(f, "%d", &c->val[k]) != 1) goto bad;
    }
    fclose(f);
    if (build_tables(c)) { nbo_free(c); return NULL; }
    return c;
bad:
    fclose(f);
    nbo_free(c);
    return NULL;
}

void nbo_free(nbo_code *c)
{
    if (!c) return;
    free(c->row_deg); free(c->row_ptr); free(c->col); free(c->val);
    free(c->bingf); free(c->addgf); free(c->mulgf); free(c->divgf);
    free(c->matUT); free(c->perm);
    free(c);
}

/* GaussianElimination, tools.c:151-218 (dense M x N, same pivoting and column swaps) */
int nbo_prepare_encoder(nbo_code *c)
{
    const int N = c->N, M = c->M, GF = c->GF;
    int m, n, k, m1, ind;
    if (c->matUT) return 0;
    int *U = calloc((size_t)M * N, sizeof(int));
    int *P = calloc(N, sizeof(int));
    for (n = 0; n < N; n++) P[n] = n;
    for (m = 0; m < M; m++)
        for (k = c->row_ptr[m]; k < c->row_ptr[m + 1]; k++) U[(size_t)m * N + c->col[k]] = c->val[k];
    for (m = 0; m < M; m++) {
        for (ind = m; ind < N; ind++) if (U[(size_t)m * N + ind] != 0) break;
        if (ind == N) { free(U); free(P); return -1; }             /* rank deficient: tools.c:181-185 */
        int t = P[ind]; P[ind] = P[m]; P[m] = t;
        for (m1 = 0; m1 < M; m1++) {
            t = U[(size_t)m1 * N + m]; U[(size_t)m1 * N + m] = U[(size_t)m1 * N + ind]; U[(size_t)m1 * N + ind] = t;
        }
        const int *rowm = U + (size_t)m * N;
        const int piv = rowm[m];
        for (m1 = m + 1; m1 < M; m1++) {
            int *r = U + (size_t)m1 * N;
            int lead = r[m];
            if (lead == 0) continue;
            for (n = m; n < N; n++) {
                int v = r[n];
                if (v != 0) v = c->divgf[v * GF + lead];             /* tools.c:201-205 */
                if (v != 0) v = c->mulgf[v * GF + piv];              /* tools.c:206-210 */
                r[n] = c->addgf[v * GF + rowm[n]];                   /* tools.c:211-214 */
            }
        }
    }
    c->matUT = U; c->perm = P;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * drand48: X(n+1) = (a X(n) + c) mod 2^48, a = 0x5DEECE66D, c = 0xB, returns X/2^48.
 * The reference never seeds it (NB_LDPC.c:88-89 only calls srand), so the stream starts from glibc's
 * zero-initialised state X0 = 0 (NOT the SVID default 0x1234ABCD330E: glibc only installs a and c lazily).
 * ---------------------------------------------------------------------------------------------- */
#define LCG_A 0x5DEECE66DULL
#define LCG_C 0xBULL
#define LCG_MASK ((1ULL << 48) - 1)
void nbo_rng_default(nbo_rng *r) { r->x = 0; }   /* glibc: static drand48 state is all-zero until seeded */
double nbo_drand48(nbo_rng *r)
{
    r->x = (LCG_A * r->x + LCG_C) & LCG_MASK;
    return ldexp((double)r->x, -48);
}
void nbo_rng_skip(nbo_rng *r, uint64_t n)
{
    uint64_t a = LCG_A, cc = LCG_C, A = 1, C = 0;          /* compose x -> A x + C by squaring */
    while (n) {
        if (n & 1) { A = (A * a) & LCG_MASK; C = (C * a + cc) & LCG_MASK; }
        cc = ((a + 1) * cc) & LCG_MASK; a = (a * a) & LCG_MASK;
        n >>= 1;
    }
    r->x = (A * r->x + C) & LCG_MASK;
}
static float my_drand48(nbo_rng *r) { return (float)nbo_drand48(r); }     /* tools.c:40,73 */

/* RandomBinaryGenerator + Encoding, tools.c:124-136, 232-268 */
void nbo_random_codeword(const nbo_code *c, nbo_rng *r, int *codeword, int *nbin)
{
    const int N = c->N, M = c->M, K = c->K, lg = c->logGF, GF = c->GF;
    int k, q, n, m;
    int *ns = calloc(N, sizeof(int));
    for (k = 0; k < K; k++) {
        int bits = 0;
        for (q = 0; q < lg; q++) {
            int b = (int)floor(my_drand48(r) * 1.9999);              /* tools.c:132 */
            bits |= b << q;
        }
        /* Bin2GF (tools.c:95): first symbol whose binary image matches */
        int s;
        for (s = 0; s < GF; s++) {
            int img = 0;
            for (q = 0; q < lg; q++) img |= c->bingf[s * lg + q] << q;
            if (img == bits) break;
        }
        ns[M + k] = s;
    }
    for (m = M - 1; m >= 0; m--) {                                    /* back-substitution, tools.c:244-254 */
        int buf = 0;
        const int *row = c->matUT + (size_t)m * N;
        for (n = m + 1; n < N; n++)
            if (row[n] != 0) buf = c->addgf[buf * GF + c->mulgf[row[n] * GF + ns[n]]];
        ns[m] = c->divgf[buf * GF + row[m]];
    }
    for (n = 0; n < N; n++) codeword[c->perm[n]] = ns[n];            /* tools.c:257-258 */
    for (n = 0; n < N; n++) for (q = 0; q < lg; q++) nbin[n * lg + q] = c->bingf[codeword[n] * lg + q];
    free(ns);
}

/* channel.c:51 */
float nbo_sigma(const nbo_code *c, float EbN)
{
    return (float)sqrt(1.0 / (2.0 * c->rate * pow(10, EbN / 10.0)));
}

/* channel.c:52-62: BPSK + Box-Muller noise, evaluated in double, stored as float */
void nbo_channel_noise(const nbo_code *c, nbo_rng *r, const int *nbin, float EbN, float *noisy)
{
    const double PI = 3.1415926536;                                    /* channel.c:18 */
    float sigma = nbo_sigma(c, EbN);
    int n, q;
    for (n = 0; n < c->N; n++)
        for (q = 0; q < c->logGF; q++) {
            float u = my_drand48(r);
            float v = my_drand48(r);
            int bpsk = 1 - 2 * nbin[n * c->logGF + q];
            noisy[n * c->logGF + q] = (float)(bpsk + sigma * sqrt(-2.0 * log(u)) * cos(2.0 * PI * v));
        }
}

/* channel.c:66-76: LLR(g) = sum_q (y_q - s_q(g))^2 / (2 sigma^2), f32 square, f64 divide+accumulate */
void nbo_channel_llr(const nbo_code *c, const float *noisy, float sigma, float *llr)
{
    const int GF = c->GF, lg = c->logGF;
    int n, g, q;
    const double den = 2.0 * (double)(float)(sigma * sigma);
    for (n = 0; n < c->N; n++)
        for (g = 0; g < GF; g++) {
            float acc = 0.0f;
            for (q = 0; q < lg; q++) {
                float d = noisy[n * lg + q] - (float)(1 - 2 * c->bingf[g * lg + q]);
                float sq = d * d;
                acc = (float)((double)acc + (double)sq / den);
            }
            llr[(size_t)n * GF + g] = acc;
        }
}

/* ---- 64-APSK channel, ModelChannel_AWGN_64 (channel.c:112-312), GF(64) only ----
 * Constellation (channel.c:133-198): DVB-S2X 8+16+20+20 APSK.  Restated from its structure: ring radius and angle
 * PI * num / den per label; the reference's initialiser evaluates radius * cos(...) in double and stores float. */
static const unsigned char apsk_ring[64] = {   /* 0: r = 1, 1: r = 2.2, 2: r = 3.6, 3: r = 5.2 */
    1,1,1,1,3,3,3,3, 1,1,1,1,2,2,2,2, 3,3,3,3,3,3,3,3, 2,2,2,2,2,2,2,2,
    0,0,0,0,3,3,3,3, 1,1,1,1,2,2,2,2, 0,0,0,0,3,3,3,3, 1,1,1,1,2,2,2,2 };
static const unsigned char apsk_num[64] = {
    25,23,7,9,7,5,1,3, 27,21,5,11,7,5,1,3, 31,29,9,11,33,27,7,13, 31,29,9,11,33,27,7,13,
    13,11,3,5,37,23,3,17, 29,19,3,13,37,23,3,17, 15,9,1,7,39,21,1,19, 31,17,1,15,39,21,1,19 };
static const unsigned char apsk_den[64] = {
    16,16,16,16,4,4,4,4, 16,16,16,16,4,4,4,4, 20,20,20,20,20,20,20,20, 20,20,20,20,20,20,20,20,
    8,8,8,8,20,20,20,20, 16,16,16,16,20,20,20,20, 8,8,8,8,20,20,20,20, 16,16,16,16,20,20,20,20 };
void nbo_apsk64_table(float *mod)
{
    static const double radius[4] = { 1.0, 2.2, 3.6, 5.2 };
    const double PI = 3.1415926536;                                    /* channel.c:18 */
    float pts[64][2], norm = 0.0f;
    int i;
    for (i = 0; i < 64; i++) {
        const double ang = PI * apsk_num[i] / apsk_den[i];
        pts[i][0] = (float)(radius[apsk_ring[i]] * cos(ang));
        pts[i][1] = (float)(radius[apsk_ring[i]] * sin(ang));
    }
    for (i = 0; i < 64; i++) norm = pts[i][0] * pts[i][0] + pts[i][1] * pts[i][1] + norm;     /* :205-211 */
    norm = sqrt(64 / norm);
    for (i = 0; i < 64; i++) { mod[2 * i] = norm * pts[i][0]; mod[2 * i + 1] = norm * pts[i][1]; }
}
float nbo_sigma_apsk64(float EbN) { return (float)sqrt(1.0 / (2.0 * pow(10, EbN / 10.0))); }      /* :232 */
/* channel.c:234-263: noisy[N][2] */
void nbo_channel_noise_apsk64(const nbo_code *c, nbo_rng *r, const int *nbin, float EbN, float *noisy)
{
    const double PI = 3.1415926536;
    const float sigma = nbo_sigma_apsk64(EbN);
    float mod[128];
    int n, q;
    nbo_apsk64_table(mod);
    for (n = 0; n < c->N; n++) {
        int som = 0;
        for (q = 0; q < 6; q++) som = som + nbin[n * 6 + q] * (1 << q);
        for (q = 0; q < 2; q++) {
            float u = my_drand48(r);
            float v = my_drand48(r);
            noisy[2 * n + q] = (float)(mod[2 * som + q] + sigma * sqrt(-2.0 * log(u)) * cos(2.0 * PI * v));
        }
    }
}
/* channel.c:266-291: TMP[k] = (I - mI)^2 / (2 sigma^2) + (Q - mQ)^2 / (2 sigma^2), f32 squares, f64 divisions and sum */
void nbo_channel_llr_apsk64(const nbo_code *c, const float *noisy, float sigma, float *llr)
{
    float mod[128];
    int n, k, q;
    nbo_apsk64_table(mod);
    for (n = 0; n < c->N; n++)
        for (k = 0; k < 64; k++) {
            int som = 0;
            float d0, d1;
            for (q = 0; q < 6; q++) som = som + c->bingf[k * 6 + q] * (1 << q);
            d0 = noisy[2 * n] - mod[2 * som]; d1 = noisy[2 * n + 1] - mod[2 * som + 1];
            llr[(size_t)n * 64 + k] = (float)((double)(d0 * d0) / (2.0 * (double)(float)(sigma * sigma)) +
                                              (double)(d1 * d1) / (2.0 * (double)(float)(sigma * sigma)));
        }
}

/* channel.c:78-91: full selection sort of one symbol's LLR vector (strict <, init 1e5 / -1) */
void nbo_sort_intrinsic(const nbo_code *c, const float *llr, float *illr, int *igf)
{
    const int GF = c->GF;
    int n, k, g;
    float tmp[256];
    for (n = 0; n < c->N; n++) {
        memcpy(tmp, llr + (size_t)n * GF, sizeof(float) * GF);
        for (k = 0; k < GF; k++) {
            float best = SENT; int bg = -1;
            for (g = 0; g < GF; g++) if (tmp[g] < best) { best = tmp[g]; bg = g; }
            illr[(size_t)n * GF + k] = best; igf[(size_t)n * GF + k] = bg;
            if (bg >= 0) tmp[bg] = SENT;      /* the reference would index -1 here; cannot happen for finite LLRs */
        }
    }
}

/* NB_LDPC.c:354-374: n_m rounds of strict-min selection, then normalisation by the first value */
void nbo_select_nm(const float *row, int GF, int n_m, float *out_llr, int *out_gf)
{
    float tmp[256];
    int k, g;
    memcpy(tmp, row, sizeof(float) * GF);
    for (k = 0; k < n_m; k++) {
        float best = SENT; int bg = 0;
        for (g = 0; g < GF; g++) if (tmp[g] < best) { best = tmp[g]; bg = g; }
        out_llr[k] = best; out_gf[k] = bg;
        tmp[bg] = SENT;
    }
    for (k = 1; k < n_m; k++) out_llr[k] = out_llr[k] - out_llr[0];
    out_llr[0] = 0.0f;
}

/* bubble_decoder.c:316-593.  Eight bubbles: p<4 walks row p to the right, p>=4 walks column p-4
 * downwards starting at row 4; candidate values are in1[i]+in2[j] (one f32 add). */
void nbo_elementary_step(const float *in1, const float *in2, const int *idx1, const int *idx2,
                         float *out, int *idxout, const int *addgf, int GF, int n_m, int nb_oper)
{
    float lo[64]; int li[64];
    float a[64], b[64]; int ia[64], ib[64];
    unsigned char seen[256];
    int bi[8], bj[8]; float bv[8];
    int p, s = 0, ss, k;
    memcpy(a, in1, sizeof(float) * n_m); memcpy(b, in2, sizeof(float) * n_m);     /* outputs may alias inputs */
    memcpy(ia, idx1, sizeof(int) * n_m); memcpy(ib, idx2, sizeof(int) * n_m);
    memset(seen, 0, sizeof seen);
    for (k = 0; k < n_m; k++) { lo[k] = SENT; li[k] = -1; }                       /* :370-374 */
    for (p = 0; p < 4; p++) { bi[p] = p; bj[p] = 0; bi[p + 4] = 4; bj[p + 4] = p; } /* :445-460 */
    for (p = 0; p < 8; p++) bv[p] = a[bi[p]] + b[bj[p]];
    for (ss = 0; ss < nb_oper; ss++) {
        float best = SENT; int pos = 0;
        for (p = 0; p < 8; p++) if (bv[p] < best) { best = bv[p]; pos = p; }     /* minimum(), :38-56 */
        int i = bi[pos], j = bj[pos];
        if (ia[i] == -1 || ib[j] == -1) break;                                    /* :478-484 */
        int g = addgf[ia[i] * GF + ib[j]];
        if (!seen[g]) { lo[s] = bv[pos]; li[s] = g; seen[g] = 1; s++; }           /* :490-496 */
        if (s == n_m) break;                                                      /* :502 */
        if (i >= n_m - 1 || j >= n_m - 1) break;                                  /* :506-544 */
        if (pos > 3) bi[pos]++; else bj[pos]++;                                   /* :548-555 */
        bv[pos] = a[bi[pos]] + b[bj[pos]];
    }
    for (k = 0; k < n_m; k++) { out[k] = lo[k]; idxout[k] = li[k]; }
}

/* bubble_decoder.c:72-305 */
void nbo_check_node_bubble(const nbo_code *c, int node, const float *vllr, const int *vgf,
                           float *cllr, int *cgf, int n_m, int nb_oper, float offset)
{
    const int GF = c->GF, dc = c->row_deg[node];
    const int *h = c->val + c->row_ptr[node];
    enum { DCMAX = 32 };
    float U[DCMAX][64]; int UI[DCMAX][64];
    float F[64], B[64]; int FI[64], BI[64];
    float inter[2 * DCMAX][64]; int interI[2 * DCMAX][64];
    float O[DCMAX][64]; int OI[DCMAX][64];
    int t, k, kk;
    for (t = 0; t < dc; t++) for (k = 0; k < n_m; k++) {                          /* rotate in, :133-152 */
        int g = vgf[t * n_m + k];
        U[t][k] = vllr[t * n_m + k];
        UI[t][k] = (g != -1) ? c->mulgf[g * GF + h[t]] : -1;
    }
    for (k = 0; k < 2 * (dc - 2); k++) for (t = 0; t < n_m; t++) { inter[k][t] = SENT; interI[k][t] = -1; }
    memcpy(F, U[0], sizeof F); memcpy(FI, UI[0], sizeof FI);
    memcpy(B, U[dc - 1], sizeof B); memcpy(BI, UI[dc - 1], sizeof BI);
    for (kk = 1; kk < dc - 1; kk++) {                                             /* :166-200 */
        memcpy(inter[kk - 1], F, sizeof F); memcpy(interI[kk - 1], FI, sizeof FI);
        memcpy(inter[2 * (dc - 2) - kk], B, sizeof B); memcpy(interI[2 * (dc - 2) - kk], BI, sizeof BI);
        nbo_elementary_step(F, U[kk], FI, UI[kk], F, FI, c->addgf, GF, n_m, nb_oper);
        nbo_elementary_step(B, U[dc - 1 - kk], BI, UI[dc - 1 - kk], B, BI, c->addgf, GF, n_m, nb_oper);
    }
    memcpy(O[dc - 1], F, sizeof F); memcpy(OI[dc - 1], FI, sizeof FI);           /* :204-213 */
    memcpy(O[0], B, sizeof B); memcpy(OI[0], BI, sizeof BI);
    for (k = 0; k < dc - 2; k++)                                                  /* :217-227 */
        nbo_elementary_step(inter[k], inter[(dc - 2) + k], interI[k], interI[(dc - 2) + k],
                            O[k + 1], OI[k + 1], c->addgf, GF, n_m, nb_oper);
    for (t = 0; t < dc; t++) {                                                    /* :231-281 */
        int stp = n_m;
        for (k = 0; k < n_m; k++) if (OI[t][k] == -1) { stp = k; break; }
        float sat = (stp > 0 ? O[t][stp - 1] : SENT) + offset;                    /* stp==0 is UB in the reference; unreachable */
        for (k = 0; k < GF; k++) { cllr[t * GF + k] = sat; cgf[t * GF + k] = k; }
        for (k = 0; k < stp; k++) cllr[t * GF + c->divgf[OI[t][k] * GF + h[t]]] = O[t][k];
    }
}

/* tools.c:312-330 */
void nbo_decision(const float *app, int N, int GF, int *decide)
{
    int n, g;
    for (n = 0; n < N; n++) {
        float best = SENT; int ind = 0;
        for (g = 0; g < GF; g++) if (app[(size_t)n * GF + g] < best) { best = app[(size_t)n * GF + g]; ind = g; }
        decide[n] = ind;
    }
}

/* tools.c:284-299: running GF sum over rows, stops at the first row that leaves it non-zero */
int nbo_syndrome(const nbo_code *c, const int *decide)
{
    int m, k, synd = 0;
    for (m = 0; m < c->M; m++) {
        for (k = c->row_ptr[m]; k < c->row_ptr[m + 1]; k++)
            synd = c->addgf[synd * c->GF + c->mulgf[c->val[k] * c->GF + decide[c->col[k]]]];
        if (synd != 0) break;
    }
    return synd;
}

/* NB_LDPC.c:266-474 */
int nbo_decode_frame(const nbo_code *c, const nbo_params *p, const float *llr,
                     int *decide, int *synd_out, int *iters_out,
                     int *decide_trace, int *synd_trace, float *app_out, float *ctov_out)
{
    const int N = c->N, GF = c->GF, E = c->E, n_m = p->n_m;
    float *APP = malloc(sizeof(float) * (size_t)N * GF);
    float *CtoV = calloc((size_t)E * GF, sizeof(float));                         /* :273-279 */
    float *mvc = malloc(sizeof(float) * (size_t)c->dc_max * GF);
    float *cl = malloc(sizeof(float) * (size_t)c->dc_max * GF);
    int *cg = malloc(sizeof(int) * (size_t)c->dc_max * GF);
    float *vl = malloc(sizeof(float) * (size_t)c->dc_max * n_m);
    int *vg = malloc(sizeof(int) * (size_t)c->dc_max * n_m);
    float *mcv = malloc(sizeof(float) * (size_t)(GF + 1));
    int iter, node, i, g, synd = 0, passes = 0;
    memcpy(APP, llr, sizeof(float) * (size_t)N * GF);                            /* :281-288 (un-sort) */
    for (iter = 0; iter < p->nb_iter_max - 1; iter++) {                          /* :314 */
        for (node = 0; node < c->M; node++) {
            const int dc = c->row_deg[node], e0 = c->row_ptr[node];
            for (i = 0; i < dc; i++) {
                const float *a = APP + (size_t)c->col[e0 + i] * GF;
                const float *cv = CtoV + (size_t)(e0 + i) * GF;
                for (g = 0; g < GF; g++) mvc[i * GF + g] = a[g] - cv[g];           /* :334 */
                nbo_select_nm(mvc + i * GF, GF, n_m, vl + i * n_m, vg + i * n_m);  /* :354-374 */
            }
            if (p->ecn == 1)
                nbo_check_node_syndrome(c, node, vl, vg, cl, cg, n_m, p->cfg, p->cfg_size, p->offset, p->n_cv);
            else
                nbo_check_node_bubble(c, node, vl, vg, cl, cg, n_m, p->nb_oper, p->offset);   /* :392 */
            for (i = 0; i < dc; i++) {
                float *a = APP + (size_t)c->col[e0 + i] * GF;
                float *cv = CtoV + (size_t)(e0 + i) * GF;
                for (g = 0; g < GF; g++) mcv[cg[i * GF + g]] = cl[i * GF + g];     /* :419 */
                for (g = 0; g < GF; g++) cv[g] = mcv[g];                          /* :438 */
                for (g = 0; g < GF; g++) a[g] = mcv[g] + mvc[i * GF + g];         /* :448 */
            }
        }
        nbo_decision(APP, N, GF, decide);                                          /* :468 */
        synd = nbo_syndrome(c, decide);                                            /* :469 */
        if (decide_trace) memcpy(decide_trace + (size_t)iter * N, decide, sizeof(int) * N);
        if (synd_trace) synd_trace[iter] = synd;
        passes++;
        if (synd == 0 && !p->force_passes) break;                                  /* :470 */
    }
    if (synd_out) *synd_out = synd;
    if (iters_out) *iters_out = iter + 1;                                          /* :474 */
    if (app_out) memcpy(app_out, APP, sizeof(float) * (size_t)N * GF);
    if (ctov_out) memcpy(ctov_out, CtoV, sizeof(float) * (size_t)E * GF);
    free(APP); free(CtoV); free(mvc); free(cl); free(cg); free(vl); free(vg); free(mcv);
    return passes;
}

/* NB_LDPC.c:250-511 */
void nbo_monte_carlo(nbo_code *c, const nbo_params *p, int nb_monte_carlo, float EbN, long *stats)
{
    const int N = c->N, GF = c->GF, lg = c->logGF;
    nbo_rng rng;
    nbo_rng_default(&rng);
    nbo_prepare_encoder(c);
    int *cw = malloc(sizeof(int) * N), *nbin = malloc(sizeof(int) * N * lg), *decide = malloc(sizeof(int) * N);
    float *noisy = malloc(sizeof(float) * N * lg), *llr = malloc(sizeof(float) * (size_t)N * GF);
    long nb, err_frames = 0, undetected = 0, bit_errors = 0, sum_it = 0;
    float sigma = nbo_sigma(c, EbN);
    for (nb = 1; nb <= nb_monte_carlo; nb++) {
        int synd, iters, k, l, e = 0;
        nbo_random_codeword(c, &rng, cw, nbin);
        nbo_channel_noise(c, &rng, nbin, EbN, noisy);
        nbo_channel_llr(c, noisy, sigma, llr);
        nbo_decode_frame(c, p, llr, decide, &synd, &iters, NULL, NULL, NULL, NULL);
        sum_it += iters;
        for (k = 0; k < c->K; k++) for (l = 0; l < lg; l++)                       /* :479-485 */
            if (c->bingf[decide[k] * lg + l] != nbin[k * lg + l]) e++;
        bit_errors += e;
        if (e) { err_frames++; if (synd == 0) undetected++; }
        if (err_frames == 40) break;                                               /* :506 */
    }
    stats[0] = (nb > nb_monte_carlo) ? nb_monte_carlo : nb;   /* frames shown on the console line */
    stats[1] = err_frames; stats[2] = undetected; stats[3] = bit_errors; stats[4] = sum_it;
    stats[5] = nb;                                            /* 'nb' after the loop: results file, :576 */
    free(cw); free(nbin); free(decide); free(noisy); free(llr);
}

/* ---- syndrome path: implemented in nbldpc_oracle_synd.c (included below when present) ---- */
#ifdef NBO_HAVE_SYND
#include "nbldpc_oracle_synd.c"
#else
int *nbo_build_config_table(int dc, int d1, int d2, int d3, int trunc, int *size_out)
{ (void)dc; (void)d1; (void)d2; (void)d3; (void)trunc; *size_out = 0; return NULL; }
void nbo_check_node_syndrome(const nbo_code *c, int node, const float *vllr, const int *vgf,
                             float *cllr, int *cgf, int n_m, const int *cfg, int cfg_size,
                             float offset, int n_cv)
{ (void)c; (void)node; (void)vllr; (void)vgf; (void)cllr; (void)cgf; (void)n_m; (void)cfg; (void)cfg_size; (void)offset; (void)n_cv; abort(); }
#endif
