/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE (oracle side), not part of the product.
 *
 * Flat-array entry points around the UNMODIFIED reference functions so that tests can call
 * LoadCode / LoadTables / ElementaryStep / CheckPassLogEMS / ModelChannel_AWGN_BPSK / Decision /
 * Syndrom / Encoding / syndrome_ems directly (ctypes cannot comfortably build int** structs).
 * Nothing here restates an algorithm: every function only marshals arrays and calls the reference.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "struct.h"
#include "init.h"
#include "tools.h"
#include "channel.h"
#include "bubble_decoder.h"
#include "syndrome_decoder.h"

void LoadCode_KN(char *FileMatrix, code_t *code);

typedef struct {
    code_t code;
    table_t table;
    decoder_t decoder;
    int have_ut;
    int **cfg;      /* syndrome config table */
    int cfg_size;
} refshim_t;

refshim_t *refshim_open(const char *path, int kn_dialect, int nbMax, int want_encoder)
{
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    fclose(f);
    refshim_t *h = calloc(1, sizeof *h);
    char *p = strdup(path);
    if (kn_dialect) LoadCode_KN(p, &h->code); else LoadCode(p, &h->code);
    free(p);
    LoadTables(&h->table, h->code.GF, h->code.logGF);
    h->decoder.nbMax = nbMax;
    AllocateDecoder(&h->code, &h->decoder);
    if (want_encoder) { GaussianElimination(&h->code, &h->table); h->have_ut = 1; }
    return h;
}

void refshim_info(const refshim_t *h, int *out8)
{
    out8[0] = h->code.N; out8[1] = h->code.M; out8[2] = h->code.GF; out8[3] = h->code.logGF;
    out8[4] = h->code.nbBranch; out8[5] = h->decoder.nbMax; out8[6] = h->code.rowDegree[0]; out8[7] = h->code.K;
}
float refshim_rate(const refshim_t *h) { return h->code.rate; }

/* flat copies of the graph: row_deg[M], col[E], val[E] in edge order */
void refshim_graph(const refshim_t *h, int *row_deg, int *col, int *val)
{
    int m, k, e = 0;
    for (m = 0; m < h->code.M; m++) {
        row_deg[m] = h->code.rowDegree[m];
        for (k = 0; k < h->code.rowDegree[m]; k++, e++) { col[e] = h->code.mat[m][k]; val[e] = h->code.matValue[m][k]; }
    }
}
/* flat copies of the tables: bingf[GF*logGF], add/mul/div[GF*GF] */
void refshim_tables(const refshim_t *h, int *bingf, int *add, int *mul, int *dv)
{
    int GF = h->code.GF, lg = h->code.logGF;
    memcpy(bingf, h->table.BINGF[0], sizeof(int) * GF * lg);
    memcpy(add, h->table.ADDGF[0], sizeof(int) * GF * GF);
    memcpy(mul, h->table.MULGF[0], sizeof(int) * GF * GF);
    memcpy(dv, h->table.DIVGF[0], sizeof(int) * GF * GF);
}

void refshim_seed_default(void)
{
    /* glibc's initial drand48 state (never seeded by the reference, NB_LDPC.c:88-89): X = 0 */
    unsigned short s[3] = { 0, 0, 0 };
    seed48(s);
}
void refshim_seed48(unsigned short s0, unsigned short s1, unsigned short s2)
{
    unsigned short s[3] = { s0, s1, s2 };
    seed48(s);
}

/* RandomBinaryGenerator + Encoding (tools.c:124, 232): outputs codeword[N], nbin[N*logGF] */
int refshim_random_codeword(refshim_t *h, int *codeword, int *nbin)
{
    if (!h->have_ut) return -1;
    int N = h->code.N, K = h->code.K, lg = h->code.logGF, n, q, idum = -1;
    int **KBIN = calloc(K, sizeof(int *)), **NBIN = calloc(N, sizeof(int *));
    int *KSYMB = calloc(K, sizeof(int));
    for (n = 0; n < K; n++) KBIN[n] = calloc(lg, sizeof(int));
    for (n = 0; n < N; n++) NBIN[n] = calloc(lg, sizeof(int));
    RandomBinaryGenerator(N, h->code.M, h->code.GF, lg, KBIN, KSYMB, h->table.BINGF, &idum);
    Encoding(&h->code, &h->table, codeword, NBIN, KSYMB);
    for (n = 0; n < N; n++) for (q = 0; q < lg; q++) nbin[n * lg + q] = NBIN[n][q];
    for (n = 0; n < K; n++) free(KBIN[n]);
    for (n = 0; n < N; n++) free(NBIN[n]);
    free(KBIN); free(NBIN); free(KSYMB);
    return 0;
}

/* ModelChannel_AWGN_BPSK (channel.c:38): outputs sorted intrinsic LLR/GF [N*GF] */
void refshim_channel_bpsk(refshim_t *h, const int *nbin, float EbN, float *illr, int *igf)
{
    int N = h->code.N, lg = h->code.logGF, GF = h->code.GF, n, q, idum = -1;
    int **NBIN = calloc(N, sizeof(int *));
    for (n = 0; n < N; n++) { NBIN[n] = calloc(lg, sizeof(int)); for (q = 0; q < lg; q++) NBIN[n][q] = nbin[n * lg + q]; }
    ModelChannel_AWGN_BPSK(&h->code, &h->decoder, &h->table, NBIN, EbN, &idum);
    memcpy(illr, h->decoder.intrinsic_LLR[0], sizeof(float) * N * GF);
    memcpy(igf, h->decoder.intrinsic_GF[0], sizeof(int) * N * GF);
    for (n = 0; n < N; n++) free(NBIN[n]);
    free(NBIN);
}

/* ModelChannel_AWGN_64 (channel.c:112): outputs sorted intrinsic LLR/GF [N*GF]; draws from the process-wide drand48 state */
void refshim_channel_apsk64(refshim_t *h, const int *nbin, float EbN, float *illr, int *igf)
{
    int N = h->code.N, lg = h->code.logGF, GF = h->code.GF, n, q, idum = -1;
    int **NBIN = calloc(N, sizeof(int *));
    for (n = 0; n < N; n++) { NBIN[n] = calloc(lg, sizeof(int)); for (q = 0; q < lg; q++) NBIN[n][q] = nbin[n * lg + q]; }
    ModelChannel_AWGN_64(&h->code, &h->decoder, NBIN, EbN, &idum);
    memcpy(illr, h->decoder.intrinsic_LLR[0], sizeof(float) * N * GF);
    memcpy(igf, h->decoder.intrinsic_GF[0], sizeof(int) * N * GF);
    for (n = 0; n < N; n++) free(NBIN[n]);
    free(NBIN);
}

/* ElementaryStep (bubble_decoder.c:316) */
void refshim_elementary_step(refshim_t *h, const float *in1, const float *in2, const int *idx1, const int *idx2,
                             float *out, int *idxout, int nbMax, int nbOper)
{
    float a[64], b[64]; int ia[64], ib[64];
    memcpy(a, in1, sizeof(float) * nbMax); memcpy(b, in2, sizeof(float) * nbMax);
    memcpy(ia, idx1, sizeof(int) * nbMax); memcpy(ib, idx2, sizeof(int) * nbMax);
    ElementaryStep(a, b, ia, ib, out, idxout, h->table.ADDGF, h->code.GF, nbMax, nbOper);
}

/* CheckPassLogEMS (bubble_decoder.c:72): in [dc*nbMax], out [dc*GF] */
void refshim_check_node(refshim_t *h, int node, const float *vllr, const int *vgf, float *cllr, int *cgf,
                        int nbOper, float offset)
{
    int dc = h->code.rowDegree[node], nm = h->decoder.nbMax, GF = h->code.GF, t, k;
    for (t = 0; t < dc; t++) for (k = 0; k < nm; k++) {
        h->decoder.M_VtoC_LLR[t][k] = vllr[t * nm + k];
        h->decoder.M_VtoC_GF[t][k] = vgf[t * nm + k];
    }
    CheckPassLogEMS(node, &h->decoder, &h->code, &h->table, nbOper, offset);
    for (t = 0; t < dc; t++) for (k = 0; k < GF; k++) {
        cllr[t * GF + k] = h->decoder.M_CtoV_LLR[t][k];
        cgf[t * GF + k] = h->decoder.M_CtoV_GF[t][k];
    }
}

/* Decision + Syndrom (tools.c:312, 284) on a flat APP[N*GF] */
int refshim_decision_syndrome(refshim_t *h, const float *app, int *decide)
{
    int N = h->code.N, GF = h->code.GF;
    memcpy(h->decoder.APP[0], app, sizeof(float) * N * GF);
    Decision(decide, h->decoder.APP, N, GF);
    return Syndrom(&h->code, decide, &h->table);
}

/* syndrome path: build_config_table + sort_config_table (+ truncation as in NB_LDPC.c:198-201) */
int refshim_build_config(refshim_t *h, int dc_max, int d1, int d2, int d3, int trunc)
{
    int size = 0;
    h->cfg = build_config_table(&size, dc_max, d1, d2, d3);
    sort_config_table(h->cfg, size, dc_max);
    if (trunc > 0 && trunc < size) size = trunc;
    h->cfg_size = size;
    return size;
}
void refshim_get_config(const refshim_t *h, int dc_max, int *out)
{
    int i, j;
    for (i = 0; i < h->cfg_size; i++) for (j = 0; j < dc_max; j++) out[i * dc_max + j] = h->cfg[i][j];
}
/* syndrome_ems (syndrome_decoder.c:26) */
int refshim_syndrome_ems(refshim_t *h, int node, const float *vllr, const int *vgf, float *cllr, int *cgf,
                         int dc_max, float offset, int n_cv)
{
    int nm = h->decoder.nbMax, GF = h->code.GF, t, k;
    for (t = 0; t < dc_max; t++) for (k = 0; k < nm; k++) {
        h->decoder.M_VtoC_LLR[t][k] = vllr[t * nm + k];
        h->decoder.M_VtoC_GF[t][k] = vgf[t * nm + k];
    }
    int r = syndrome_ems(node, &h->decoder, &h->code, &h->table, h->cfg, h->cfg_size, dc_max, offset, n_cv);
    for (t = 0; t < dc_max; t++) for (k = 0; k < GF; k++) {
        cllr[t * GF + k] = h->decoder.M_CtoV_LLR[t][k];
        cgf[t * GF + k] = h->decoder.M_CtoV_GF[t][k];
    }
    return r;
}
