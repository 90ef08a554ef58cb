/*
 * oracle/ref_probes.c -- TEST INFRASTRUCTURE (oracle side), not part of the product.
 *
 * Interposition probes linked with the UNMODIFIED reference main() (NB_LDPC.c compiled with
 * -Dmain=ref_main -DDecision=probe_Decision ... see oracle/Makefile).  Every probe calls the real
 * reference function and records its inputs/outputs, so the reference's own decode loop
 * (NB_LDPC.c:250-511) runs untouched and becomes the golden-trace generator and the CPU baseline.
 *
 * Environment:
 *   NBREF_TRACE=<file>   write a tagged binary trace (format below)
 *   NBREF_LEVEL=<0..3>   0 summary only, 1 +per-pass decide/synd (+intrinsic LLR), 2 +APP per pass,
 *                        3 +every check-node call's inputs/outputs
 *   NBREF_FORCE=1        Syndrom reports non-zero while passes remain -> fixed number of passes
 *                        (the true syndrome is still traced)
 *   NBREF_DIALECT=kn     parse the matrix with the reference's KN branch (init.c:211-227)
 *   NBREF_SUMMARY=<file> append one line "frames decode_s channel_s passes" at exit
 *   NBREF_ECN=syndrome   the check node is the reference's syndrome_ems (syndrome_decoder.c:26) instead of
 *                        CheckPassLogEMS -- i.e. NB_LDPC.c:388 un-commented and :392 commented -- with the
 *                        table of NB_LDPC.c:189-201 built by the reference's own build_config_table +
 *                        sort_config_table.  NBREF_SYND="d1,d2,d3,trunc,ncv" (default 19,15,5,1000,25; the
 *                        values in the commented code, d_1=40, overrun the n_m-wide message rows)
 *
 * Trace record: int32 tag, int32 payload_bytes, payload.
 *   tag 1 HEADER  int32[8]  N M GF logGF nbMax nbBranch dc0 NbOper
 *   tag 2 FRAME   int8 [N*logGF]    codeword bits NBIN (channel.c:38 input)
 *   tag 3 INTRIN  float[N*GF] LLR, then int16[N*GF] symbols   (level>=1)
 *   tag 4 PASS    int32 synd_true, then int16[N] decide        (level>=1)
 *   tag 5 APP     float[N*GF]                                   (level>=2)
 *   tag 6 CN      int32 node, int32 dc, float in_llr[dc*nbMax], int16 in_gf[dc*nbMax],
 *                 float out_llr[dc*GF], int16 out_gf[dc*GF]     (level>=3)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "struct.h"
#include "init.h"
#include "tools.h"
#include "channel.h"
#include "bubble_decoder.h"
#include "syndrome_decoder.h"

void LoadCode_KN(char *FileMatrix, code_t *code);
int ref_main(int argc, char *argv[]);

static FILE *g_tr;
static int g_level, g_force, g_hdr_done;
static int g_N, g_GF, g_nbMax, g_NbOper;
static long g_frames, g_passes;
static double g_decode_s, g_channel_s;
static struct timespec g_t0;
static int g_in_frame;
static code_t *g_code;

static double now_diff(const struct timespec *a)
{
    struct timespec b;
    clock_gettime(CLOCK_MONOTONIC, &b);
    return (double)(b.tv_sec - a->tv_sec) + 1e-9 * (double)(b.tv_nsec - a->tv_nsec);
}

static void rec(int tag, const void *p, int nbytes)
{
    if (!g_tr) return;
    fwrite(&tag, 4, 1, g_tr);
    fwrite(&nbytes, 4, 1, g_tr);
    if (nbytes) fwrite(p, 1, (size_t)nbytes, g_tr);
}

static void probes_init(void)
{
    static int done;
    if (done) return;
    done = 1;
    const char *t = getenv("NBREF_TRACE");
    if (t && *t) g_tr = fopen(t, "wb");
    const char *l = getenv("NBREF_LEVEL");
    g_level = l ? atoi(l) : 1;
    const char *f = getenv("NBREF_FORCE");
    g_force = f ? atoi(f) : 0;
}

void probe_LoadCode(char *FileMatrix, code_t *code)
{
    probes_init();
    const char *d = getenv("NBREF_DIALECT");
    if (d && strcmp(d, "kn") == 0) LoadCode_KN(FileMatrix, code);
    else LoadCode(FileMatrix, code);
}

void probe_Channel(code_t *code, decoder_t *decoder, table_t *table, int **NBIN, float EbN, int *init_rand)
{
    probes_init();
    struct timespec c0;
    clock_gettime(CLOCK_MONOTONIC, &c0);
    ModelChannel_AWGN_BPSK(code, decoder, table, NBIN, EbN, init_rand);
    g_channel_s += now_diff(&c0);
    g_code = code;
    g_N = code->N; g_GF = code->GF; g_nbMax = decoder->nbMax;
    if (g_tr) {
        if (!g_hdr_done) {
            int h[8] = { code->N, code->M, code->GF, code->logGF, decoder->nbMax, code->nbBranch,
                         code->rowDegree[0], g_NbOper };
            rec(1, h, sizeof h);
            g_hdr_done = 1;
        }
        int n, q, nb = code->N * code->logGF;
        signed char *bits = malloc((size_t)nb);
        for (n = 0; n < code->N; n++)
            for (q = 0; q < code->logGF; q++) bits[n * code->logGF + q] = (signed char)NBIN[n][q];
        rec(2, bits, nb);
        free(bits);
        if (g_level >= 1) {
            int cnt = code->N * code->GF, i;
            int tag = 3, bytes = cnt * 6;
            short *s = malloc((size_t)cnt * 2);
            for (i = 0; i < cnt; i++) s[i] = (short)decoder->intrinsic_GF[0][i];
            fwrite(&tag, 4, 1, g_tr); fwrite(&bytes, 4, 1, g_tr);
            fwrite(decoder->intrinsic_LLR[0], 4, (size_t)cnt, g_tr);
            fwrite(s, 2, (size_t)cnt, g_tr);
            free(s);
        }
    }
    g_frames++;
    g_in_frame = 1;
    clock_gettime(CLOCK_MONOTONIC, &g_t0);   /* decode time starts when the channel returns */
}

static int g_synd = -1, g_cfg_size, g_ncv;
static int **g_cfg;
static void check_node(int node, decoder_t *decoder, code_t *code, table_t *table, int NbOper, float offset)
{
    if (g_synd < 0) {
        const char *e = getenv("NBREF_ECN");
        g_synd = (e && strcmp(e, "syndrome") == 0);
        if (g_synd) {
            int d1 = 19, d2 = 15, d3 = 5, trunc = 1000, dc = code->rowDegree[0];
            g_ncv = 25;
            const char *p = getenv("NBREF_SYND");
            if (p) sscanf(p, "%d,%d,%d,%d,%d", &d1, &d2, &d3, &trunc, &g_ncv);
            g_cfg = build_config_table(&g_cfg_size, dc, d1, d2, d3);
            sort_config_table(g_cfg, g_cfg_size, dc);
            if (trunc > 0 && trunc < g_cfg_size) g_cfg_size = trunc;
        }
    }
    if (g_synd) syndrome_ems(node, decoder, code, table, g_cfg, g_cfg_size, code->rowDegree[node], offset, g_ncv);
    else CheckPassLogEMS(node, decoder, code, table, NbOper, offset);
}

void probe_CheckPass(int node, decoder_t *decoder, code_t *code, table_t *table, int NbOper, float offset)
{
    g_NbOper = NbOper;
    if (g_tr && g_level >= 3) {
        int dc = code->rowDegree[node], nm = decoder->nbMax, GF = code->GF, t, k;
        float *il = malloc(sizeof(float) * dc * nm);
        short *ig = malloc(sizeof(short) * dc * nm);
        for (t = 0; t < dc; t++) for (k = 0; k < nm; k++) {
            il[t * nm + k] = decoder->M_VtoC_LLR[t][k];
            ig[t * nm + k] = (short)decoder->M_VtoC_GF[t][k];
        }
        check_node(node, decoder, code, table, NbOper, offset);
        short *og = malloc(sizeof(short) * dc * GF);
        float *ol = malloc(sizeof(float) * dc * GF);
        for (t = 0; t < dc; t++) for (k = 0; k < GF; k++) {
            ol[t * GF + k] = decoder->M_CtoV_LLR[t][k];
            og[t * GF + k] = (short)decoder->M_CtoV_GF[t][k];
        }
        int tag = 6, bytes = 8 + dc * nm * 6 + dc * GF * 6;
        fwrite(&tag, 4, 1, g_tr); fwrite(&bytes, 4, 1, g_tr);
        fwrite(&node, 4, 1, g_tr); fwrite(&dc, 4, 1, g_tr);
        fwrite(il, 4, (size_t)dc * nm, g_tr); fwrite(ig, 2, (size_t)dc * nm, g_tr);
        fwrite(ol, 4, (size_t)dc * GF, g_tr); fwrite(og, 2, (size_t)dc * GF, g_tr);
        free(il); free(ig); free(ol); free(og);
        return;
    }
    check_node(node, decoder, code, table, NbOper, offset);
}

static float **g_app;
void probe_Decision(int *decision, float **APP, int N, int GF)
{
    g_app = APP;
    Decision(decision, APP, N, GF);
}

int probe_Syndrom(code_t *code, int *decide, table_t *tableGF)
{
    int synd = Syndrom(code, decide, tableGF);
    g_passes++;
    if (g_in_frame) { g_decode_s += now_diff(&g_t0); }
    if (g_tr && g_level >= 1) {
        int n, tag = 4, bytes = 4 + 2 * code->N;
        short *d = malloc((size_t)code->N * 2);
        for (n = 0; n < code->N; n++) d[n] = (short)decide[n];
        fwrite(&tag, 4, 1, g_tr); fwrite(&bytes, 4, 1, g_tr);
        fwrite(&synd, 4, 1, g_tr); fwrite(d, 2, (size_t)code->N, g_tr);
        free(d);
        if (g_level >= 2 && g_app) rec(5, g_app[0], code->N * code->GF * 4);
    }
    if (g_in_frame) clock_gettime(CLOCK_MONOTONIC, &g_t0);
    if (g_force && synd == 0) return 1;   /* keep iterating; NB_LDPC.c:470 only tests for zero */
    return synd;
}

__attribute__((destructor)) static void probes_fini(void)
{
    if (g_tr) fclose(g_tr);
    const char *s = getenv("NBREF_SUMMARY");
    if (s && *s) {
        FILE *f = fopen(s, "a");
        if (f) {
            fprintf(f, "%ld %.9f %.9f %ld\n", g_frames, g_decode_s, g_channel_s, g_passes);
            fclose(f);
        }
    }
}

int main(int argc, char *argv[])
{
    probes_init();
    return ref_main(argc, argv);
}
