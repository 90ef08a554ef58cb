/*
 * nbldpc_b200.h -- C ABI of the B200-native EMS NB-LDPC decoder (libnbldpc_b200.so).
 *
 * Drop-in boundary for the decode path of Lcrypto/EMS-decoder-of-NB-LDPC-codes.  The reference has no
 * FFI; its "interface" for this path is the set of C calls made by main() (NB_LDPC.c).  Every entry
 * point below names the reference call(s) it replaces.  Plain pointers and sizes only; no CUDA, torch
 * or C++ types appear here.  All functions return 0 (NBGPU_OK) or a negative NBGPU_E* code and never
 * call exit(); the message is available from nbgpu_last_error().  A context is bound to one device
 * and one CUDA stream; different contexts are independent.  There is NO CPU fallback: every compute
 * entry point fails with NBGPU_ECUDA when no sm_100 device is usable.
 *
 * Conventions kept from the reference:
 *   - LLRs are float32 "distances": smaller = more likely; +1e5 = absent.
 *   - GF symbols are ints 0..q-1: 0 = zero element, k = alpha^(k-1); q in {16,64,256}.
 *   - edge e of check node m is the e-th entry in file order (numB in NB_LDPC.c:266-460).
 *   - "nb_iter_max" is argv[2]: at most nb_iter_max-1 decoding passes run (NB_LDPC.c:314).
 */
#ifndef NBLDPC_B200_H
#define NBLDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBGPU_OK        0
#define NBGPU_EINVAL   -1   /* bad argument / unsupported parameter combination */
#define NBGPU_EIO      -2   /* matrix file missing or malformed                  */
#define NBGPU_ENOMEM   -3
#define NBGPU_ECUDA    -4   /* CUDA error or no usable GPU                        */
#define NBGPU_ESTATE   -5   /* requested frame state no longer resident           */
#define NBGPU_ERANK    -6   /* H is rank deficient (tools.c:181-185)             */

/* ------------------------------------------------------------------------------------------------
 * Host side (plain C, no GPU needed): code description = code_t + table_t of the reference.
 * ---------------------------------------------------------------------------------------------- */
typedef struct nbgpu_code nbgpu_code;

#define NBGPU_ALIST_AUTO 0
#define NBGPU_ALIST_UBS  1   /* init.c:195-207 : all columns (0-based) then all coefficients (1..q-1) */
#define NBGPU_ALIST_KN   2   /* init.c:211-227 : per row (col 1-based, exponent 0..q-2) pairs          */
#define NBGPU_ALIST_FULL 3   /* full alist (max-degree line, column lists, then row lists of (col 1-based, exponent) pairs):
                                matrices/KN/N64800_*, which the reference's LoadCode cannot read */

/* replaces LoadCode (init.c:143) + LoadTables (init.c:427).  The alist dialect is a runtime switch
 * instead of the compile-time '#define KN_matrix' (init.c:25). */
int nbgpu_code_load(nbgpu_code **out, const char *path, int dialect);

/* Bind an already loaded reference code_t/table_t (flattened by the caller):
 * row_deg[M]; col[E], val[E] in edge order; bingf[q*logq]; addgf/mulgf/divgf[q*q] (may be NULL: the
 * tables are then generated exactly as init.c:37-130 does). */
int nbgpu_code_from_arrays(nbgpu_code **out, int N, int M, int q, const int *row_deg, const int *col,
                           const int *val, const int *bingf, const int *addgf, const int *mulgf,
                           const int *divgf);
void nbgpu_code_free(nbgpu_code *c);

/* info[10] = N, M, K, q, logq, E, dc_max, dc_min, dialect used, reserved */
void nbgpu_code_info(const nbgpu_code *c, int *info);
float nbgpu_code_rate(const nbgpu_code *c);                                  /* code_t.rate, init.c:167 */
/* copies of code_t.mat/matValue (edge order) and of the table_t tables, as ints like the reference */
void nbgpu_code_graph(const nbgpu_code *c, int *row_deg, int *col, int *val);
void nbgpu_code_tables(const nbgpu_code *c, int *bingf, int *addgf, int *mulgf, int *divgf);

/* Frame source on the host, replaces RandomBinaryGenerator + GaussianElimination + Encoding
 * (tools.c:124, 151, 232) and the noise part of ModelChannel_AWGN_BPSK (channel.c:51-62).
 * The generator is glibc's drand48 recurrence; state = 48-bit X (the reference never seeds it, so its
 * stream starts from X = 0).  All frames of a Monte-Carlo run can be generated independently with
 * nbgpu_rng_skip. */
typedef struct { uint64_t x; } nbgpu_rng;
void nbgpu_rng_reference_default(nbgpu_rng *r);
void nbgpu_rng_skip(nbgpu_rng *r, uint64_t ndraws);
double nbgpu_rng_drand48(nbgpu_rng *r);
int nbgpu_code_prepare_encoder(nbgpu_code *c);                               /* tools.c:151; NBGPU_ERANK */
int nbgpu_random_codeword(const nbgpu_code *c, nbgpu_rng *r, int *codeword /*[N]*/, int *nbin /*[N*logq]*/);
float nbgpu_sigma(const nbgpu_code *c, float EbN);                           /* channel.c:51 */
int nbgpu_awgn_bpsk_noise(const nbgpu_code *c, nbgpu_rng *r, const int *nbin, float EbN,
                          float *noisy /*[N*logq]*/);                        /* channel.c:52-62 */

/* ---- 64-APSK over AWGN: ModelChannel_AWGN_64, channel.c:112-312 (GF(64) codes only) ----
 * nbgpu_apsk64_table     the normalised DVB-S2X 8+16+20+20 APSK constellation, mod[64][2], indexed by the symbol's binary
 *                        image (channel.c:133-222)
 * nbgpu_sigma_apsk64     sigma = sqrt(1 / (2 * 10^(EbN/10))), channel.c:232
 * nbgpu_awgn_apsk64_noise  noisy[N][2] = constellation point + Box-Muller noise on the caller's drand48 stream, same draw
 *                        order as the reference (per symbol: u, v for I, then u, v for Q), channel.c:234-263 */
int nbgpu_apsk64_table(float *mod /*[64][2]*/);
float nbgpu_sigma_apsk64(float EbN);
int nbgpu_awgn_apsk64_noise(const nbgpu_code *c, nbgpu_rng *r, const int *nbin /*[N][6] or NULL = all-zero word*/, float EbN,
                            float *noisy /*[N][2]*/);


/* Configuration table of the syndrome-based check node: replaces build_config_table + sort_config_table
 * (syndrome_decoder.c:1542, 2285) and the truncation NB_LDPC.c:198-201.  Returns the number of
 * configurations (rows of dc ints) or a negative error; table may be NULL to query the size. */
int nbgpu_config_table(int dc, int d1, int d2, int d3, int trunc, int *table, int capacity);

/* ------------------------------------------------------------------------------------------------
 * Device side
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int   n_m;          /* decoder.nbMax, argv[5]; 5 <= n_m <= min(q,32)                          */
    int   nb_oper;      /* NbOper, argv[7]: max pops per ElementaryStep                            */
    int   nb_iter_max;  /* NbIterMax, argv[2] (>= 2)                                               */
    float offset;       /* argv[6]                                                                 */
    int   ecn_kind;     /* 0 = CheckPassLogEMS (L-Bubble F/B), 1 = syndrome_ems                    */
    int   early_stop;   /* 1 = stop a frame at the first zero syndrome (NB_LDPC.c:470); 0 = always
                           run nb_iter_max-1 passes (fixed-iteration throughput mode)               */
    /* syndrome_ems only (NB_LDPC.c:185-201, commented out there; see DESIGN.md "Syndrome path").
       0 = default: d1 = n_m-1, d2 = min(15, n_m-1), d3 = min(5, n_m-1), cfg_trunc = 1000, n_cv = nb_oper
       (NB_LDPC.c:185), border = 4 (syndrome_decoder.c:56; the only supported value) */
    int   d1, d2, d3, cfg_trunc, n_cv, border;
    /* tuning; 0 = automatic */
    int   frames_per_cta;  /* frames decoded together by one CTA                                  */
    int   cns_per_step;    /* upper bound on the check nodes (over all frames of a CTA) per step  */
} nbgpu_params;

typedef struct nbgpu_ctx nbgpu_ctx;

/* replaces AllocateDecoder (init.c:310): owns all device memory; code is only read during the call */
int nbgpu_create(nbgpu_ctx **out, const nbgpu_code *code, const nbgpu_params *p, int device, int max_batch);
void nbgpu_destroy(nbgpu_ctx *ctx);                                          /* FreeDecoder, init.c:395 */
const char *nbgpu_last_error(const nbgpu_ctx *ctx);   /* ctx may be NULL: error of the last failed create/load */

/* Intake (a13 minus the RNG): LLR part of ModelChannel_AWGN_BPSK (channel.c:66-91) for B frames.
 * noisy[B][N][logq] are the received BPSK samples (NoisyBin).  Outputs (host, any may be NULL):
 * llr[B][N][q] dense in GF order (what NB_LDPC.c:281-288 scatters into APP), and the sorted
 * decoder_t.intrinsic_LLR / intrinsic_GF [B][N][q]. */
int nbgpu_channel_awgn_bpsk(nbgpu_ctx *ctx, const float *noisy, float sigma, int B,
                            float *llr, float *intrinsic_llr, int *intrinsic_gf);

/* Replaces the per-frame loop body NB_LDPC.c:266-474 for B frames (B <= max_batch).
 * Inputs are HOST buffers; outputs (host): decide[B][N] hard decisions (tools.c:312), synd[B] value
 * of the last Syndrom call (tools.c:284; 0 <=> codeword), iters[B] = iter+1 as added to sum_it
 * (NB_LDPC.c:474). */
int nbgpu_decode_noisy(nbgpu_ctx *ctx, const float *noisy /*[B][N][logq]*/, float sigma, int B,
                       int *decide, int *synd, int *iters);
int nbgpu_decode_llr(nbgpu_ctx *ctx, const float *llr /*[B][N][q] GF order*/, int B,
                     int *decide, int *synd, int *iters);
/* LLR part of ModelChannel_AWGN_64 on the GPU (channel.c:266-308): dense LLR[B][N][64] and, when illr/igf are given, the sorted
 * intrinsic_LLR / intrinsic_GF; nbgpu_decode_apsk64 = the same intake fused into the decoder (replaces channel.c:266-308 +
 * NB_LDPC.c:266-474 for B frames); nbgpu_upload_apsk64 + nbgpu_run + nbgpu_download is the resident-input form. */
int nbgpu_channel_awgn_apsk64(nbgpu_ctx *ctx, const float *noisy /*[B][N][2]*/, float sigma, int B,
                              float *llr, float *illr, int *igf);
int nbgpu_decode_apsk64(nbgpu_ctx *ctx, const float *noisy /*[B][N][2]*/, float sigma, int B,
                        int *decide, int *synd, int *iters);

/* Same work split into stages so that a caller can keep inputs resident in HBM (bench.py "value"):
 * upload (H2D) -> run (kernel only, asynchronous on the ctx stream) -> download (D2H + sync). */
int nbgpu_upload_noisy(nbgpu_ctx *ctx, const float *noisy, float sigma, int B);
int nbgpu_upload_apsk64(nbgpu_ctx *ctx, const float *noisy /*[B][N][2]*/, float sigma, int B);   /* 64-APSK samples, see below */
int nbgpu_upload_llr(nbgpu_ctx *ctx, const float *llr, int B);
int nbgpu_run(nbgpu_ctx *ctx);                  /* decode the resident batch; returns after launch    */
int nbgpu_sync(nbgpu_ctx *ctx);
int nbgpu_download(nbgpu_ctx *ctx, int *decide, int *synd, int *iters);
/* device time (CUDA events on the ctx stream) of the decode kernel in the last nbgpu_run, ms */
int nbgpu_last_kernel_ms(nbgpu_ctx *ctx, float *ms);
/* number of kernels launched by this ctx since creation (bench.py gpu_launches) */
long nbgpu_launch_count(const nbgpu_ctx *ctx);
/* device-side stopwatch on the ctx stream (CUDA events): begin records, end records + waits + returns ms */
int nbgpu_timer_begin(nbgpu_ctx *ctx);
int nbgpu_timer_end(nbgpu_ctx *ctx, float *ms);
/* page-lock / unlock a caller-owned host buffer so that the H2D/D2H copies of nbgpu_decode_* are DMA copies */
int nbgpu_host_register(void *ptr, size_t bytes);
int nbgpu_host_unregister(void *ptr);
/* geometry chosen by nbgpu_create: geo[8] = grid (CTAs), frames per CTA group, check nodes per step and frame,
 * steps per pass, dynamic shared memory bytes, resident frame slots, warps per CTA, check nodes per warp */
int nbgpu_geometry(const nbgpu_ctx *ctx, int *geo);
/* selection rows that needed the exact (slow) scan since creation; diagnostic */
long nbgpu_slow_selects(nbgpu_ctx *ctx);

/* Frame source on the DEVICE: frames [frame0, frame0 + B) of the reference's Monte-Carlo stream are
 * produced where the decoder reads them, so a simulation never moves a frame across PCIe.  Replaces, per frame,
 * RandomBinaryGenerator (tools.c:124-136), Encoding (tools.c:232-270; the elimination of tools.c:151-218 is done once
 * on the host by nbgpu_code_prepare_encoder) and the noise of ModelChannel_AWGN_BPSK (channel.c:51-62).
 * `origin` is the drand48 state before frame 0 (nbgpu_rng_reference_default for the reference's stream).
 * The samples are bit-identical to nbgpu_random_codeword + nbgpu_awgn_bpsk_noise: the few samples whose f32 rounding
 * could depend on the last bits of the device's log/cos are detected and recomputed by the host's libm
 * (nbgpu_source_fixups tells how many in the last call).  Afterwards the batch is resident as after
 * nbgpu_upload_noisy: call nbgpu_run, then nbgpu_source_results (or nbgpu_download for the decisions). */
int  nbgpu_source_frames(nbgpu_ctx *ctx, nbgpu_code *code, const nbgpu_rng *origin, uint64_t frame0, int B, float EbN);
/* what the source produced (either may be NULL): codeword symbols [B][N] (CodeWord of tools.c:258), noisy [B][N][logq] */
int  nbgpu_source_download(nbgpu_ctx *ctx, int *codeword, float *noisy);
/* after nbgpu_run: information-bit errors of each frame (NB_LDPC.c:479-485), syndrome value, iteration count;
 * feed them to the frame-order statistics rule (NB_LDPC.c:474-507).  12 bytes per frame cross PCIe. */
int  nbgpu_source_results(nbgpu_ctx *ctx, int *bit_errors, int *synd, int *iters);
long nbgpu_source_fixups(const nbgpu_ctx *ctx);
/* test hook: relative width of the "ambiguous rounding" test (default 2^-46, i.e. 64 ulp of a double) */
int  nbgpu_source_set_margin(nbgpu_ctx *ctx, double margin);

/* Parity/debug: APP[N][q] and CtoV[E][q] (dense, as decoder_t.APP / decoder_t.CtoV) of one frame of
 * the last batch.  Only frames whose working set is still resident can be read (NBGPU_ESTATE). */
int nbgpu_get_state(nbgpu_ctx *ctx, int frame, float *APP, float *CtoV);

/* Boundary 1 (unit parity): one check node for B independent input sets.
 * replaces CheckPassLogEMS (bubble_decoder.h:17) / syndrome_ems (syndrome_decoder.h:14):
 * vllr/vgf[B][dc][n_m] = decoder->M_VtoC_LLR/GF before the call, cllr/cgf[B][dc][q] =
 * decoder->M_CtoV_LLR/GF after it. */
int nbgpu_check_node(nbgpu_ctx *ctx, int node, const float *vllr, const int *vgf,
                     float *cllr, int *cgf, int B);
/* ElementaryStep (bubble_decoder.h:28) for B independent pairs of n_m-lists (symbols, -1 = absent) */
int nbgpu_elementary_step(nbgpu_ctx *ctx, const float *in1, const float *in2, const int *idx1,
                          const int *idx2, float *out, int *idxout, int B);
/* the inline truncation NB_LDPC.c:354-374: rows[B][q] -> llr/gf[B][n_m] */
int nbgpu_select_nm(nbgpu_ctx *ctx, const float *rows, float *llr, int *gf, int B);
/* Decision + Syndrom (tools.c:312, 284) on dense APP[B][N][q] */
int nbgpu_decision_syndrome(nbgpu_ctx *ctx, const float *app, int *decide, int *synd, int B);

/* Monte-Carlo statistics of NB_LDPC.c:474-507 for frames decoded in frame order.
 * codeword_bits[B][N][logq]; stats[6] accumulates: frames, erroneous frames, undetected, bit errors,
 * sum_it, stop flag (1 once the 40th erroneous frame was reached; later frames are ignored). */
int nbgpu_accumulate_stats(const nbgpu_code *c, const int *codeword_bits, const int *decide,
                           const int *synd, const int *iters, int B, long *stats);
/* the same rule fed with the per-frame error counts of nbgpu_source_results */
int nbgpu_accumulate_results(const int *bit_errors, const int *synd, const int *iters, int B, long *stats);

/* build information */
const char *nbgpu_version(void);
int nbgpu_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NBLDPC_B200_H */
