/*
 * oracle/nbldpc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the reference's EMS NB-LDPC decode path
 * (Lcrypto/EMS-decoder-of-NB-LDPC-codes).  It exists to CHECK the CUDA product; nothing under
 * ems-decoder-of-nb-ldpc-codes_b200/ links, imports or executes it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the oracle
 * is pinned against the reference itself, compiled here from /root/reference by oracle/Makefile into
 * oracle/_ref (function-level via libref.so, whole-loop via essai_probe traces) and against the
 * committed traces in tests/golden/ that were generated from that build (tests/golden/make_golden.py).
 */
#ifndef NBLDPC_ORACLE_H
#define NBLDPC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int N, M, K, GF, logGF, E, dc_max;
    float rate;          /* (float)(N-M)/N, init.c:167 */
    int *row_deg;        /* [M]                     code_t.rowDegree   */
    int *row_ptr;        /* [M+1] first edge of row (running numB, NB_LDPC.c:460) */
    int *col;            /* [E]  code_t.mat, edge order                */
    int *val;            /* [E]  code_t.matValue, edge order           */
    int *bingf;          /* [GF*logGF]  table_t.BINGF                  */
    int *addgf, *mulgf, *divgf;   /* [GF*GF]  table_t.ADDGF/MULGF/DIVGF */
    int *matUT;          /* [M*N] or NULL: upper-triangular H (tools.c:151) */
    int *perm;           /* [N]   or NULL                               */
} nbo_code;

typedef struct {
    int n_m;             /* decoder.nbMax (argv[5])  */
    int nb_oper;         /* NbOper (argv[7])         */
    int nb_iter_max;     /* NbIterMax (argv[2]); passes <= nb_iter_max-1, NB_LDPC.c:314 */
    float offset;        /* argv[6]                  */
    int ecn;             /* 0 = CheckPassLogEMS (bubble), 1 = syndrome_ems */
    int force_passes;    /* 1 = ignore synd==0 (fixed number of passes, for timing) */
    /* syndrome path (ecn==1) */
    int n_cv;
    int cfg_size;
    const int *cfg;      /* [cfg_size*dc] */
} nbo_params;

typedef struct { uint64_t x; } nbo_rng;   /* glibc drand48 state (48 bits) */

/* dialect: 1 = UBS alist (init.c:195-207), 2 = KN alist (init.c:211-227) */
nbo_code *nbo_load(const char *path, int dialect);
void nbo_free(nbo_code *c);
int nbo_prepare_encoder(nbo_code *c);                       /* GaussianElimination, tools.c:151 */

void nbo_rng_default(nbo_rng *r);                           /* glibc's unseeded state            */
void nbo_rng_skip(nbo_rng *r, uint64_t ndraws);             /* LCG jump-ahead                    */
double nbo_drand48(nbo_rng *r);

void nbo_random_codeword(const nbo_code *c, nbo_rng *r, int *codeword, int *nbin);  /* tools.c:124,232 */
float nbo_sigma(const nbo_code *c, float EbN);                                         /* channel.c:51  */
void nbo_channel_noise(const nbo_code *c, nbo_rng *r, const int *nbin, float EbN, float *noisy); /* channel.c:52-62 */
void nbo_channel_llr(const nbo_code *c, const float *noisy, float sigma, float *llr);  /* channel.c:66-76 */
void nbo_sort_intrinsic(const nbo_code *c, const float *llr, float *illr, int *igf);   /* channel.c:78-91 */
/* 64-APSK channel (ModelChannel_AWGN_64, channel.c:112-312), GF(64) codes */
void nbo_apsk64_table(float *mod /*[64][2]*/);                                         /* channel.c:133-222 */
float nbo_sigma_apsk64(float EbN);                                                     /* channel.c:232 */
void nbo_channel_noise_apsk64(const nbo_code *c, nbo_rng *r, const int *nbin, float EbN, float *noisy /*[N][2]*/);   /* :234-263 */
void nbo_channel_llr_apsk64(const nbo_code *c, const float *noisy, float sigma, float *llr);                        /* :266-291 */

void nbo_select_nm(const float *row, int GF, int n_m, float *out_llr, int *out_gf);    /* NB_LDPC.c:354-374 */
void nbo_elementary_step(const float *in1, const float *in2, const int *idx1, const int *idx2,
                         float *out, int *idxout, const int *addgf, int GF, int n_m, int nb_oper); /* bubble_decoder.c:316 */
void nbo_check_node_bubble(const nbo_code *c, int node, const float *vllr, const int *vgf,
                           float *cllr, int *cgf, int n_m, int nb_oper, float offset);  /* bubble_decoder.c:72 */
void nbo_decision(const float *app, int N, int GF, int *decide);                       /* tools.c:312 */
int nbo_syndrome(const nbo_code *c, const int *decide);                                /* tools.c:284 */

/* One frame of NB_LDPC.c:266-474.  llr = dense channel LLR [N*GF] in GF order.
 * Outputs: decide[N], *synd (last Syndrom value), *iters (= iter+1 as accumulated at NB_LDPC.c:474).
 * Optional traces (may be NULL): decide_trace[(nb_iter_max-1)*N], synd_trace[nb_iter_max-1],
 * app_out[N*GF] (APP after the last executed pass), ctov_out[E*GF].  Returns passes executed. */
int nbo_decode_frame(const nbo_code *c, const nbo_params *p, const float *llr,
                     int *decide, int *synd, int *iters,
                     int *decide_trace, int *synd_trace, float *app_out, float *ctov_out);

/* Whole Monte-Carlo run of NB_LDPC.c:250-511; stats[6] = frames run (nb), erroneous frames,
 * undetected, bit errors, sum_it, frames counted in the results file (nb after the loop). */
void nbo_monte_carlo(nbo_code *c, const nbo_params *p, int nb_monte_carlo, float EbN, long *stats);

/* syndrome path */
int *nbo_build_config_table(int dc, int d1, int d2, int d3, int trunc, int *size_out);  /* syndrome_decoder.c:1542,1661,2285 */
void nbo_check_node_syndrome(const nbo_code *c, int node, const float *vllr, const int *vgf,
                             float *cllr, int *cgf, int n_m, const int *cfg, int cfg_size,
                             float offset, int n_cv);                                     /* syndrome_decoder.c:26 */

#ifdef __cplusplus
}
#endif
#endif
