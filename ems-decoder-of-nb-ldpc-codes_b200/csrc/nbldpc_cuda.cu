/*
 * nbldpc_cuda.cu -- CUDA layer + C ABI of the B200 EMS NB-LDPC decoder (sm_100a).
 *
 * One persistent CTA decodes one GROUP of F frames at a time, start to finish (all passes, early
 * termination included), pulling groups from an atomic work queue; the grid is a multiple of the SM
 * count.  A decoding pass walks the host-built step schedule (nbldpc_host.c): every step holds
 * check nodes that share no variable, so a CTA-wide barrier between steps reproduces the reference's
 * sequential layered update (NB_LDPC.c:320-466) bit for bit.  One step = three phases:
 *
 *   phase 1  one WARP per edge   Mvc = APP - CtoV (NB_LDPC.c:334), top-n_m selection + normalisation
 *                                (:354-374), rotation by the edge coefficient (bubble_decoder.c:133)
 *   phase 2  one THREAD per ElementaryStep: forward/backward chains, then the merges
 *                                (bubble_decoder.c:157-227, 316-593)
 *   phase 3  one WARP per edge   saturation + offset, expansion to the dense q-vector
 *                                (bubble_decoder.c:231-281), CtoV store, APP = Mcv + Mvc
 *                                (NB_LDPC.c:415-450), fused Decision (tools.c:312)
 *
 * HBM layout per resident frame (slot): APP[N][q] f32 (row = 4q bytes, one coalesced warp access);
 * CtoV as one lossless record per edge {llr[n_m] f32, sat f32, stp i32, sym[n_m] u8} -- a dense CtoV
 * row is "stp explicit (symbol, LLR) pairs + one constant" (bubble_decoder.c:262-270); decisions u8.
 * Between phase 1 and 3 the APP row temporarily holds the un-normalised Mvc (NB_LDPC.c:448 needs it).
 */
#include "nbldpc_device.cuh"
#include "nbldpc_internal.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define NT 256            /* threads per CTA */
#define NW (NT / 32)

struct KArgs {
    int N, M, E, q, logq, dc_max;
    int n_m, nb_oper, passes, early_stop, nb_iter_max;
    float offset;
    int F, G, nsteps, L, tasks;        /* frames per group, items per step, lists per item, ES tasks */
    int B, input_kind;                  /* 0 = noisy samples, 1 = dense LLR */
    double den;                         /* 2.0 * (double)(float)(sigma*sigma), channel.c:73 */
    const int *row_ptr, *col, *order, *step_ptr, *isolated;
    int n_isolated;
    const uint8_t *hval, *last, *rotin, *rotout, *img, *inv;
    const float *in;
    float *app; uint8_t *ctov; uint8_t *dec;
    int rec_stride;
    int *out_decide, *out_synd, *out_iters, *frame_slot, *slot_frame;
    unsigned *queue, *slow_counter;
    /* shared memory carve-up (byte offsets) */
    int off_llr, off_sym, off_len, off_mask, off_ws, off_misc, smem_bytes;
};

__device__ __forceinline__ float rec_sat(const uint8_t *rec, int n_m) { return *reinterpret_cast<const float *>(rec + 4 * n_m); }
__device__ __forceinline__ int rec_stp(const uint8_t *rec, int n_m) { return *reinterpret_cast<const int *>(rec + 4 * n_m + 4); }

/* dense CtoV values of this lane's symbols from the record (through ws.row) */
template <int Q>
__device__ __forceinline__ void expand_record(const uint8_t *rec, int n_m, int lane, WarpScratch<Q> &ws,
                                              float (&c)[QTraits<Q>::VPL])
{
    constexpr int VPL = QTraits<Q>::VPL;
    const int stp = rec_stp(rec, n_m);
    const float sat = rec_sat(rec, n_m);
    if (stp == 0) {
#pragma unroll
        for (int j = 0; j < VPL; j++) c[j] = sat;
        return;
    }
#pragma unroll
    for (int j = 0; j < VPL; j++) ws.row[lane * VPL + j] = sat;
    __syncwarp();
    if (lane < stp) ws.row[rec[4 * n_m + 8 + lane]] = reinterpret_cast<const float *>(rec)[lane];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < VPL; j++) c[j] = ws.row[lane * VPL + j];
    __syncwarp();
}

/* phase 3 core: from the check node's output list of one edge (binary-image symbols) build the record
 * and the dense Mcv values of this lane (bubble_decoder.c:231-281). */
template <int Q>
__device__ __forceinline__ void finish_edge(const float *lo, const uint8_t *so, int len, const uint8_t *rotout_h,
                                            int n_m, float offset, int lane, WarpScratch<Q> &ws, uint8_t *rec,
                                            float (&mcv)[QTraits<Q>::VPL])
{
    constexpr int VPL = QTraits<Q>::VPL;
    const int stp = len;                                         /* first absent entry, :233-243 */
    float llr = NB_SENT; int sym = 0;
    if (lane < stp) { llr = lo[lane]; sym = rotout_h[so[lane]]; }   /* DIVGF by the coefficient, :249-254 */
    const float last = __shfl_sync(NB_FULL, llr, max(stp - 1, 0));
    const float sat = __fadd_rn(stp > 0 ? last : NB_SENT, offset);  /* :264 (stp == 0 cannot occur) */
    if (rec) {
        if (lane < n_m) { reinterpret_cast<float *>(rec)[lane] = llr; rec[4 * n_m + 8 + lane] = (uint8_t)sym; }
        if (lane == 0) { *reinterpret_cast<float *>(rec + 4 * n_m) = sat; *reinterpret_cast<int *>(rec + 4 * n_m + 4) = stp; }
    }
#pragma unroll
    for (int j = 0; j < VPL; j++) ws.row[lane * VPL + j] = sat;
    __syncwarp();
    if (lane < stp) ws.row[sym] = llr;                           /* :267-270 */
    __syncwarp();
#pragma unroll
    for (int j = 0; j < VPL; j++) mcv[j] = ws.row[lane * VPL + j];
    __syncwarp();
}

/* list addressing inside the CTA's shared memory */
struct Lists {
    float *llr; uint8_t *sym; uint8_t *len; int L, n_m;
    __device__ __forceinline__ float *l(int item, int li) const { return llr + (item * L + li) * n_m; }
    __device__ __forceinline__ uint8_t *s(int item, int li) const { return sym + (item * L + li) * n_m; }
    __device__ __forceinline__ uint8_t &n(int item, int li) const { return len[item * L + li]; }
};
/* list ids of one item with degree dc: U[t] = t; F after s steps = dc+s-1 (s>=1); B after s steps =
 * dc+(dc-2)+s-1; merge k = dc+2(dc-2)+k  (bubble_decoder.c:166-227: MatriceInter rows) */
__device__ __forceinline__ int id_F(int dc, int s) { return s == 0 ? 0 : dc + s - 1; }
__device__ __forceinline__ int id_B(int dc, int s) { return s == 0 ? dc - 1 : dc + (dc - 2) + s - 1; }
__device__ __forceinline__ int id_M(int dc, int k) { return dc + 2 * (dc - 2) + k; }
/* output list of edge t: t=0 -> B after dc-2 steps, t=dc-1 -> F after dc-2 steps, else merge t-1 */
__device__ __forceinline__ int id_out(int dc, int t)
{
    if (t == 0) return id_B(dc, dc - 2);
    if (t == dc - 1) return id_F(dc, dc - 2);
    return id_M(dc, t - 1);
}

/* forward (dir 0) or backward (dir 1) chain of one check node: dc-2 sequential elementary steps */
__device__ __forceinline__ void chain_task(const Lists &ls, int item, int dc, int dir, uint32_t *mask, int mstride,
                                           int mwords, int nb_oper)
{
    for (int kk = 1; kk <= dc - 2; kk++) {
        const int a = dir ? id_B(dc, kk - 1) : id_F(dc, kk - 1);
        const int b = dir ? dc - 1 - kk : kk;
        const int o = dir ? id_B(dc, kk) : id_F(dc, kk);
        ls.n(item, o) = (uint8_t)es_serial(ls.l(item, a), ls.s(item, a), ls.n(item, a), ls.l(item, b), ls.s(item, b),
                                           ls.n(item, b), ls.l(item, o), ls.s(item, o), mask, mstride, mwords, ls.n_m, nb_oper);
    }
}
/* merge k: ElementaryStep(F after k steps, B after dc-3-k steps) -> output of edge k+1 */
__device__ __forceinline__ void merge_task(const Lists &ls, int item, int dc, int k, uint32_t *mask, int mstride,
                                           int mwords, int nb_oper)
{
    const int a = id_F(dc, k), b = id_B(dc, dc - 3 - k), o = id_M(dc, k);
    ls.n(item, o) = (uint8_t)es_serial(ls.l(item, a), ls.s(item, a), ls.n(item, a), ls.l(item, b), ls.s(item, b),
                                       ls.n(item, b), ls.l(item, o), ls.s(item, o), mask, mstride, mwords, ls.n_m, nb_oper);
}

/* LLR intake for one variable (channel.c:66-76), one warp: lanes < 2*logq first build the per-bit
 * terms (double)((y-s)^2) / (2 sigma^2), then every lane accumulates its symbols bit by bit with the
 * reference's float <- double + double rounding. */
template <int Q>
__device__ __forceinline__ void intake_variable(const float *noisy_n, double den, const uint8_t *img, int lane,
                                                WarpScratch<Q> &ws, float (&v)[QTraits<Q>::VPL])
{
    constexpr int VPL = QTraits<Q>::VPL;
    constexpr int LOGQ = QTraits<Q>::LOGQ;
    double *t = reinterpret_cast<double *>(ws.row);
    if (lane < 2 * LOGQ) {
        const float y = noisy_n[lane >> 1];
        const float s = (lane & 1) ? -1.0f : 1.0f;                 /* BPSK(b) = 1 - 2b */
        const float d = __fsub_rn(y, s);
        const float sq = __fmul_rn(d, d);
        t[lane] = __ddiv_rn((double)sq, den);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < VPL; j++) {
        const int g = lane * VPL + j;
        float acc = 0.0f;
        if (Q >= 32 || lane < Q) {
            const int im = img[g];
#pragma unroll
            for (int b = 0; b < LOGQ; b++) acc = __double2float_rn(__dadd_rn((double)acc, t[2 * b + ((im >> b) & 1)]));
        }
        v[j] = acc;
    }
    __syncwarp();
}

template <int Q>
__global__ void __launch_bounds__(NT, 2) decode_kernel(const KArgs a)
{
    constexpr int VPL = QTraits<Q>::VPL;
    extern __shared__ __align__(16) unsigned char smem[];
    Lists ls;
    ls.llr = reinterpret_cast<float *>(smem + a.off_llr);
    ls.sym = smem + a.off_sym;
    ls.len = smem + a.off_len;
    ls.L = a.L; ls.n_m = a.n_m;
    uint32_t *masks = reinterpret_cast<uint32_t *>(smem + a.off_mask);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpScratch<Q> &ws = reinterpret_cast<WarpScratch<Q> *>(smem + a.off_ws)[warp];
    int *misc = reinterpret_cast<int *>(smem + a.off_misc);
    int *s_base = misc;                 /* [1]  first frame of the group  */
    int *s_alive = misc + 1;            /* [1]  frames of the group still iterating */
    int *s_done = misc + 2;             /* [F]  0 = iterating, else iters value to report */
    int *s_badrow = misc + 2 + a.F;     /* [F]  first check row with non-zero syndrome */
    int *s_synd = misc + 2 + 2 * a.F;   /* [F]  last syndrome value */
    const int F = a.F, N = a.N, n_m = a.n_m, dcm = a.dc_max;
    const int mwords = (Q + 31) / 32;
    const size_t slot_app = (size_t)F * N * Q;
    float *app = a.app + blockIdx.x * slot_app;
    uint8_t *ctov = a.ctov + (size_t)blockIdx.x * F * a.E * a.rec_stride;
    uint8_t *dec = a.dec + (size_t)blockIdx.x * F * N;

    for (;;) {
        if (tid == 0) *s_base = (int)atomicAdd(a.queue, (unsigned)F);
        __syncthreads();
        const int base = *s_base;
        if (base >= a.B) break;
        const int nf = min(F, a.B - base);
        /* ---------------- frame initialisation: NB_LDPC.c:273-288 + channel.c:66-76 ---------------- */
        for (int w = warp; w < nf * N; w += NW) {
            const int f = w / N, n = w - f * N;
            float v[VPL];
            if (a.input_kind == 0) intake_variable<Q>(a.in + ((size_t)(base + f) * N + n) * a.logq, a.den, a.img, lane, ws, v);
            else load_row<Q>(a.in + ((size_t)(base + f) * N + n) * Q, lane, v);
            store_row<Q>(app + ((size_t)f * N + n) * Q, lane, v);
        }
        for (int i = tid; i < nf * a.E; i += NT) {           /* CtoV = 0: stp 0, constant 0.0f */
            uint8_t *rec = ctov + (size_t)i * a.rec_stride;
            *reinterpret_cast<float *>(rec + 4 * n_m) = 0.0f;
            *reinterpret_cast<int *>(rec + 4 * n_m + 4) = 0;
        }
        for (int i = tid; i < F; i += NT) { s_done[i] = (i < nf) ? 0 : -1; s_synd[i] = 0; }
        if (tid == 0) { *s_alive = nf; for (int f = 0; f < nf; f++) { a.frame_slot[base + f] = blockIdx.x * F + f; a.slot_frame[blockIdx.x * F + f] = base + f; } }
        __syncthreads();
        /* variables that no check node touches keep their channel decision */
        for (int w = warp; w < nf * a.n_isolated; w += NW) {
            const int f = w / a.n_isolated, n = a.isolated[w - f * a.n_isolated];
            float v[VPL];
            load_row<Q>(app + ((size_t)f * N + n) * Q, lane, v);
            const int d = warp_argmin<Q>(v, lane);
            if (lane == 0) dec[f * N + n] = (uint8_t)d;
        }

        for (int pass = 0; pass < a.passes; pass++) {
            for (int st = 0; st < a.nsteps; st++) {
                const int c0 = a.step_ptr[st], ncn = a.step_ptr[st + 1] - c0;
                const int items = ncn * nf;                   /* item = f * ncn + ci */
                /* ---------------- phase 1 ---------------- */
                for (int w = warp; w < items * dcm; w += NW) {
                    const int item = w / dcm, t = w - item * dcm;
                    const int f = item / ncn, cn = a.order[c0 + item - f * ncn];
                    const int e0 = a.row_ptr[cn], dc = a.row_ptr[cn + 1] - e0;
                    if (t >= dc || s_done[f]) continue;
                    const int e = e0 + t;
                    float *row = app + ((size_t)f * N + a.col[e]) * Q;
                    const uint8_t *rec = ctov + ((size_t)f * a.E + e) * a.rec_stride;
                    float v[VPL], c[VPL];
                    load_row<Q>(row, lane, v);
                    expand_record<Q>(rec, n_m, lane, ws, c);
#pragma unroll
                    for (int j = 0; j < VPL; j++) v[j] = __fsub_rn(v[j], c[j]);           /* NB_LDPC.c:334 */
                    store_row<Q>(row, lane, v);
                    store_row<Q>(ws.row, lane, v);
                    float llr; int sym;
                    warp_select_nm<Q>(v, lane, ws, n_m, llr, sym, a.slow_counter);
                    if (lane < n_m) {
                        ls.l(item, t)[lane] = llr;
                        ls.s(item, t)[lane] = a.rotin[a.hval[e] * Q + sym];                /* bubble_decoder.c:145 */
                    }
                    if (lane == 0) ls.n(item, t) = (uint8_t)n_m;
                }
                __syncthreads();
                /* ---------------- phase 2a: forward / backward chains ---------------- */
                for (int task = tid; task < items * 2; task += NT) {
                    const int item = task >> 1, f = item / ncn, cn = a.order[c0 + item - f * ncn];
                    const int dc = a.row_ptr[cn + 1] - a.row_ptr[cn];
                    if (s_done[f]) continue;
                    chain_task(ls, item, dc, task & 1, masks + (task % a.tasks), a.tasks, mwords, a.nb_oper);
                }
                __syncthreads();
                /* ---------------- phase 2b: merges ---------------- */
                if (dcm > 2) {
                    for (int task = tid; task < items * (dcm - 2); task += NT) {
                        const int item = task / (dcm - 2), k = task - item * (dcm - 2);
                        const int f = item / ncn, cn = a.order[c0 + item - f * ncn];
                        const int dc = a.row_ptr[cn + 1] - a.row_ptr[cn];
                        if (k >= dc - 2 || s_done[f]) continue;
                        merge_task(ls, item, dc, k, masks + (task % a.tasks), a.tasks, mwords, a.nb_oper);
                    }
                    __syncthreads();
                }
                /* ---------------- phase 3 ---------------- */
                for (int w = warp; w < items * dcm; w += NW) {
                    const int item = w / dcm, t = w - item * dcm;
                    const int f = item / ncn, cn = a.order[c0 + item - f * ncn];
                    const int e0 = a.row_ptr[cn], dc = a.row_ptr[cn + 1] - e0;
                    if (t >= dc || s_done[f]) continue;
                    const int e = e0 + t, var = a.col[e];
                    float *row = app + ((size_t)f * N + var) * Q;
                    uint8_t *rec = ctov + ((size_t)f * a.E + e) * a.rec_stride;
                    const int lo_id = id_out(dc, t);
                    float mcv[VPL], v[VPL];
                    load_row<Q>(row, lane, v);                                             /* Mvc parked by phase 1 */
                    finish_edge<Q>(ls.l(item, lo_id), ls.s(item, lo_id), ls.n(item, lo_id), a.rotout + a.hval[e] * Q,
                                   n_m, a.offset, lane, ws, rec, mcv);
#pragma unroll
                    for (int j = 0; j < VPL; j++) v[j] = __fadd_rn(mcv[j], v[j]);          /* NB_LDPC.c:448 */
                    store_row<Q>(row, lane, v);
                    if (a.last[e]) {                                                       /* tools.c:312 fused */
                        const int d = warp_argmin<Q>(v, lane);
                        if (lane == 0) dec[f * N + var] = (uint8_t)d;
                    }
                }
                __syncthreads();
            }
            /* ---------------- Syndrom (tools.c:284-299) + early termination (NB_LDPC.c:470) ---------------- */
            for (int i = tid; i < nf; i += NT) s_badrow[i] = a.M;
            __syncthreads();
            for (int i = tid; i < nf * a.M; i += NT) {
                const int f = i / a.M, m = i - f * a.M;
                if (s_done[f]) continue;
                int x = 0;
                for (int e = a.row_ptr[m]; e < a.row_ptr[m + 1]; e++) x ^= a.rotin[a.hval[e] * Q + dec[f * N + a.col[e]]];
                if (x) atomicMin(&s_badrow[f], m);
            }
            __syncthreads();
            if (tid < nf && !s_done[tid]) {
                const int f = tid, m = s_badrow[f];
                int x = 0;
                if (m < a.M) for (int e = a.row_ptr[m]; e < a.row_ptr[m + 1]; e++) x ^= a.rotin[a.hval[e] * Q + dec[f * N + a.col[e]]];
                s_synd[f] = a.inv[x];
                if ((x == 0 && a.early_stop)) { s_done[f] = pass + 1; atomicSub(s_alive, 1); }   /* sum_it += iter+1 */
            }
            __syncthreads();
            if (*s_alive == 0) break;
        }
        /* ---------------- results ---------------- */
        for (int i = tid; i < nf * N; i += NT) a.out_decide[(size_t)base * N + i] = dec[i];
        if (tid < nf) {
            a.out_synd[base + tid] = s_synd[tid];
            a.out_iters[base + tid] = s_done[tid] > 0 ? s_done[tid] : a.nb_iter_max;      /* NB_LDPC.c:474 */
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------------
 * unit-boundary kernels (parity tests call these through the C ABI)
 * ---------------------------------------------------------------------------------------------- */
template <int Q>
__global__ void select_kernel(const float *rows, float *llr, int *gf, int B, int n_m, unsigned *slow)
{
    constexpr int VPL = QTraits<Q>::VPL;
    __shared__ WarpScratch<Q> wss[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch<Q> &ws = wss[warp];
    for (int r = blockIdx.x * NW + warp; r < B; r += gridDim.x * NW) {
        float v[VPL];
        load_row<Q>(rows + (size_t)r * Q, lane, v);
        store_row<Q>(ws.row, lane, v);
        float l; int s;
        warp_select_nm<Q>(v, lane, ws, n_m, l, s, slow);
        if (lane < n_m) { llr[(size_t)r * n_m + lane] = l; gf[(size_t)r * n_m + lane] = s; }
        __syncwarp();
    }
}

/* ElementaryStep on B pairs; symbols already binary images (0..q-1) with lens */
__global__ void es_kernel(const float *in1, const float *in2, const uint8_t *s1, const uint8_t *s2, const int *len1,
                          const int *len2, float *out, uint8_t *so, int *leno, uint32_t *maskbuf, int B, int n_m,
                          int nb_oper, int mwords)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    leno[i] = es_serial(in1 + (size_t)i * n_m, s1 + (size_t)i * n_m, len1[i], in2 + (size_t)i * n_m, s2 + (size_t)i * n_m,
                        len2[i], out + (size_t)i * n_m, so + (size_t)i * n_m, maskbuf + (size_t)i * mwords, 1, mwords, n_m, nb_oper);
}

/* one check node (bubble ECN) for B input sets: same chain/merge/finish code as the decoder */
template <int Q>
__global__ void checknode_kernel(const KArgs a, int node, const float *vllr, const int *vgf, float *cllr, int *cgf, int B)
{
    constexpr int VPL = QTraits<Q>::VPL;
    extern __shared__ __align__(16) unsigned char smem[];
    Lists ls;
    ls.llr = reinterpret_cast<float *>(smem + a.off_llr);
    ls.sym = smem + a.off_sym;
    ls.len = smem + a.off_len;
    ls.L = a.L; ls.n_m = a.n_m;
    uint32_t *masks = reinterpret_cast<uint32_t *>(smem + a.off_mask);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpScratch<Q> &ws = reinterpret_cast<WarpScratch<Q> *>(smem + a.off_ws)[warp];
    const int e0 = a.row_ptr[node], dc = a.row_ptr[node + 1] - e0, n_m = a.n_m;
    const int mwords = (Q + 31) / 32;
    for (int b0 = blockIdx.x * a.G; b0 < B; b0 += gridDim.x * a.G) {
        const int items = min(a.G, B - b0);
        for (int i = tid; i < items * dc * n_m; i += NT) {
            const int item = i / (dc * n_m), r = i - item * dc * n_m, t = r / n_m, k = r - t * n_m;
            const size_t src = ((size_t)(b0 + item) * dc + t) * n_m + k;
            ls.l(item, t)[k] = vllr[src];
            const int g = vgf[src];
            ls.s(item, t)[k] = a.rotin[a.hval[e0 + t] * Q + (g & (Q - 1))];
            if (k == 0) ls.n(item, t) = (uint8_t)n_m;
        }
        __syncthreads();
        for (int task = tid; task < items * 2; task += NT)
            chain_task(ls, task >> 1, dc, task & 1, masks + (task % a.tasks), a.tasks, mwords, a.nb_oper);
        __syncthreads();
        for (int task = tid; task < items * (dc - 2); task += NT)
            merge_task(ls, task / (dc - 2), dc, task % (dc - 2), masks + (task % a.tasks), a.tasks, mwords, a.nb_oper);
        __syncthreads();
        for (int w = warp; w < items * dc; w += NW) {
            const int item = w / dc, t = w - item * dc, lo_id = id_out(dc, t);
            float mcv[VPL];
            finish_edge<Q>(ls.l(item, lo_id), ls.s(item, lo_id), ls.n(item, lo_id), a.rotout + a.hval[e0 + t] * Q, n_m,
                           a.offset, lane, ws, nullptr, mcv);
            float *dst = cllr + ((size_t)(b0 + item) * dc + t) * Q;
            store_row<Q>(dst, lane, mcv);
            int *gdst = cgf + ((size_t)(b0 + item) * dc + t) * Q;
#pragma unroll
            for (int j = 0; j < VPL; j++) if (Q >= 32 || lane < Q) gdst[lane * VPL + j] = lane * VPL + j;   /* :276 */
        }
        __syncthreads();
    }
}

/* Decision + Syndrom on dense APP[B][N][q] */
template <int Q>
__global__ void decision_kernel(const KArgs a, const float *app, int *decide, int B)
{
    constexpr int VPL = QTraits<Q>::VPL;
    const int lane = threadIdx.x & 31;
    const long total = (long)B * a.N;
    for (long r = (long)blockIdx.x * NW + (threadIdx.x >> 5); r < total; r += (long)gridDim.x * NW) {
        float v[VPL];
        load_row<Q>(app + r * Q, lane, v);
        const int d = warp_argmin<Q>(v, lane);
        if (lane == 0) decide[r] = d;
    }
}
__global__ void syndrome_kernel(const KArgs a, const int *decide, int *synd, int B)
{
    /* one CTA per frame */
    __shared__ int badrow;
    for (int f = blockIdx.x; f < B; f += gridDim.x) {
        if (threadIdx.x == 0) badrow = a.M;
        __syncthreads();
        for (int m = threadIdx.x; m < a.M; m += blockDim.x) {
            int x = 0;
            for (int e = a.row_ptr[m]; e < a.row_ptr[m + 1]; e++) x ^= a.rotin[a.hval[e] * a.q + decide[(size_t)f * a.N + a.col[e]]];
            if (x) atomicMin(&badrow, m);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int x = 0, m = badrow;
            if (m < a.M) for (int e = a.row_ptr[m]; e < a.row_ptr[m + 1]; e++) x ^= a.rotin[a.hval[e] * a.q + decide[(size_t)f * a.N + a.col[e]]];
            synd[f] = a.inv[x];
        }
        __syncthreads();
    }
}

/* channel intake as a standalone kernel: dense LLR and (optionally) the sorted intrinsic arrays
 * (channel.c:66-91).  The sort is only needed for interface parity with decoder_t.intrinsic_*. */
template <int Q>
__global__ void channel_kernel(const KArgs a, const float *noisy, float *llr, float *illr, int *igf, int B)
{
    constexpr int VPL = QTraits<Q>::VPL;
    __shared__ WarpScratch<Q> wss[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch<Q> &ws = wss[warp];
    const long total = (long)B * a.N;
    for (long r = (long)blockIdx.x * NW + warp; r < total; r += (long)gridDim.x * NW) {
        float v[VPL];
        intake_variable<Q>(noisy + r * a.logq, a.den, a.img, lane, ws, v);
        if (llr) store_row<Q>(llr + r * Q, lane, v);
        if (illr) {
            /* full stable sort = q rounds of the exact scan (channel.c:78-91) */
            float tmp[VPL];
#pragma unroll
            for (int j = 0; j < VPL; j++) tmp[j] = v[j];
            for (int k = 0; k < Q; k++) {
                float bv = NB_SENT; int bg = 0x7fffffff;
#pragma unroll
                for (int j = 0; j < VPL; j++) if ((Q >= 32 || lane < Q) && tmp[j] < bv) { bv = tmp[j]; bg = lane * VPL + j; }
                warp_lexmin(bv, bg);
#pragma unroll
                for (int j = 0; j < VPL; j++) if (bg == lane * VPL + j) tmp[j] = NB_SENT;
                if (lane == 0) { illr[r * Q + k] = bv; igf[r * Q + k] = (bg == 0x7fffffff) ? -1 : bg; }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------- */
struct nbgpu_ctx {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1, ev_t0, ev_t1;
    int per_sm;
    nbgpu_params p;
    KArgs k;
    int N, M, E, q, logq, dc_max;
    int max_batch, nslots, grid;
    int *row_ptr_h;  /* host copies for get_state */
    int *inv_h;
    /* device buffers */
    int *d_row_ptr, *d_col, *d_order, *d_step_ptr, *d_isolated;
    uint8_t *d_hval, *d_last, *d_rotin, *d_rotout, *d_img, *d_inv;
    float *d_app; uint8_t *d_ctov; uint8_t *d_dec;
    float *d_in; size_t in_capacity;
    int *d_decide, *d_synd, *d_iters, *d_frame_slot, *d_slot_frame;
    unsigned *d_queue, *d_slow;
    int resident_B, resident_kind;
    long launches;
    float last_ms;
    char err[512];
};

static void ctx_err(nbgpu_ctx *c, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    char buf[512];
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) snprintf(c->err, sizeof c->err, "%s", buf);
    nbgpu_set_global_error("%s", buf);
}
#define CK(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx_err(ctx, "%s failed: %s", #call, cudaGetErrorString(e_)); return NBGPU_ECUDA; } } while (0)

extern "C" const char *nbgpu_last_error(const nbgpu_ctx *ctx) { return ctx ? ctx->err : nbgpu_get_global_error(); }

extern "C" int nbgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <typename T> static int upload(nbgpu_ctx *c, T **dst, const std::vector<T> &src)
{
    CK(c, cudaMalloc((void **)dst, src.size() * sizeof(T) + 16));
    CK(c, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return NBGPU_OK;
}

static int align_up(int x, int a) { return (x + a - 1) / a * a; }

/* shared-memory plan for G items */
static void plan_smem(KArgs &k, int G, size_t ws_bytes)
{
    k.G = G;
    k.L = 4 * k.dc_max - 6; if (k.L < 2) k.L = 2;
    k.tasks = G * (k.dc_max - 2 > 2 ? k.dc_max - 2 : 2);
    int off = 0;
    k.off_llr = off; off += G * k.L * k.n_m * 4;
    k.off_sym = off; off += align_up(G * k.L * k.n_m, 16);
    k.off_len = off; off += align_up(G * k.L, 16);
    k.off_mask = off; off += align_up(((k.q + 31) / 32) * k.tasks * 4, 16);
    k.off_ws = off; off += (int)ws_bytes * NW;
    k.off_misc = off; off += align_up((2 + 3 * k.F) * 4, 16);
    k.smem_bytes = off;
}

template <int Q> static size_t ws_size() { return sizeof(WarpScratch<Q>); }
static size_t ws_size_q(int q) { return q == 16 ? ws_size<16>() : q == 64 ? ws_size<64>() : ws_size<256>(); }

extern "C" int nbgpu_create(nbgpu_ctx **out, const nbgpu_code *code, const nbgpu_params *p, int device, int max_batch)
{
    *out = NULL;
    if (!code || !p) { ctx_err(NULL, "nbgpu_create: NULL argument"); return NBGPU_EINVAL; }
    if (p->ecn_kind != 0) { ctx_err(NULL, "ecn_kind %d: only the L-Bubble check node (0) is built in this version", p->ecn_kind); return NBGPU_EINVAL; }
    if (p->n_m < 5 || p->n_m > 32 || p->n_m > code->q) { ctx_err(NULL, "n_m=%d unsupported: need 5 <= n_m <= min(q,32) (ElementaryStep reads row 4, bubble_decoder.c:457)", p->n_m); return NBGPU_EINVAL; }
    if (p->nb_iter_max < 2) { ctx_err(NULL, "nb_iter_max must be >= 2 (nb_iter_max-1 passes are run, NB_LDPC.c:314)"); return NBGPU_EINVAL; }
    if (p->nb_oper < 1) { ctx_err(NULL, "nb_oper must be >= 1"); return NBGPU_EINVAL; }
    if (code->dc_min < 2 || code->dc_max > 16) { ctx_err(NULL, "check degree range [%d,%d] unsupported (2..16)", code->dc_min, code->dc_max); return NBGPU_EINVAL; }
    if (max_batch < 1) { ctx_err(NULL, "max_batch must be >= 1"); return NBGPU_EINVAL; }
    for (int m = 0; m < code->M; m++)
        for (int e = code->row_ptr[m]; e < code->row_ptr[m + 1]; e++)
            for (int e2 = e + 1; e2 < code->row_ptr[m + 1]; e2++)
                if (code->col[e] == code->col[e2]) { ctx_err(NULL, "check node %d references variable %d twice", m, code->col[e]); return NBGPU_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); ctx_err(NULL, "no CUDA device available (this library has no CPU fallback)"); return NBGPU_ECUDA; }
    if (device < 0 || device >= ndev) { ctx_err(NULL, "device %d out of range (0..%d)", device, ndev - 1); return NBGPU_EINVAL; }
    nbgpu_ctx *c = (nbgpu_ctx *)calloc(1, sizeof *c);
    c->device = device; c->p = *p; c->max_batch = max_batch;
    c->N = code->N; c->M = code->M; c->E = code->E; c->q = code->q; c->logq = code->logq; c->dc_max = code->dc_max;
    CK(c, cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(c, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { ctx_err(c, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); free(c); return NBGPU_ECUDA; }
    CK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(c, cudaEventCreate(&c->ev0)); CK(c, cudaEventCreate(&c->ev1));
    CK(c, cudaEventCreate(&c->ev_t0)); CK(c, cudaEventCreate(&c->ev_t1));

    const int q = code->q, E = code->E, N = code->N, M = code->M;
    KArgs &k = c->k;
    memset(&k, 0, sizeof k);
    k.N = N; k.M = M; k.E = E; k.q = q; k.logq = code->logq; k.dc_max = code->dc_max;
    k.n_m = p->n_m; k.nb_oper = p->nb_oper; k.passes = p->nb_iter_max - 1; k.early_stop = p->early_stop;
    k.nb_iter_max = p->nb_iter_max; k.offset = p->offset;
    k.rec_stride = align_up(5 * p->n_m + 8, 16);

    /* tables */
    std::vector<uint8_t> hval(E), last(E, 0), rotin((size_t)q * q), rotout((size_t)q * q), img(q), inv(q);
    std::vector<int> lastedge(N, -1), isolated;
    for (int e = 0; e < E; e++) { hval[e] = (uint8_t)code->val[e]; lastedge[code->col[e]] = e; }
    for (int n = 0; n < N; n++) { if (lastedge[n] >= 0) last[lastedge[n]] = 1; else isolated.push_back(n); }
    for (int s = 0; s < q; s++) { img[s] = (uint8_t)code->img[s]; inv[s] = (uint8_t)code->inv[s]; }
    for (int h = 0; h < q; h++) for (int s = 0; s < q; s++) {
        rotin[(size_t)h * q + s] = (uint8_t)code->img[code->mulgf[s * q + h]];                     /* MULGF[sym][h] */
        rotout[(size_t)h * q + s] = h ? (uint8_t)code->divgf[code->inv[s] * q + h] : 0;            /* DIVGF[sym][h] */
    }
    c->row_ptr_h = (int *)malloc(sizeof(int) * (M + 1)); memcpy(c->row_ptr_h, code->row_ptr, sizeof(int) * (M + 1));
    c->inv_h = (int *)malloc(sizeof(int) * q); memcpy(c->inv_h, code->inv, sizeof(int) * q);

    /* launch geometry: shared-memory budget -> items per step G, frames per group F, step schedule */
    const int budget = getenv("NBGPU_SMEM_KB") ? atoi(getenv("NBGPU_SMEM_KB")) * 1024 : 100 * 1024;
    k.F = 1;
    int G = 64;
    if (p->cns_per_step > 0) G = p->cns_per_step;
    for (;; G--) { plan_smem(k, G, ws_size_q(q)); if (k.smem_bytes <= budget || G == 1) break; }
    /* choose F: smallest number of frames per group that keeps the steps reasonably full */
    int bestF = 1; double bestU = -1;
    nbgpu_schedule sched; memset(&sched, 0, sizeof sched);
    for (int F = 1; F <= G && F <= 64; F *= 2) {
        if (p->frames_per_cta > 0) F = p->frames_per_cta;
        nbgpu_schedule s2;
        nbgpu_build_schedule(code, G / F > 0 ? G / F : 1, &s2);
        const double util = (double)M * F / ((double)s2.nsteps * G);
        nbgpu_free_schedule(&s2);
        if (util > bestU * 1.10) { bestU = util; bestF = F; }
        if (p->frames_per_cta > 0 || util > 0.85) break;
    }
    k.F = bestF < 1 ? 1 : bestF;
    if (k.F > G) k.F = G;
    plan_smem(k, G, ws_size_q(q));
    nbgpu_build_schedule(code, G / k.F > 0 ? G / k.F : 1, &sched);
    k.nsteps = sched.nsteps;
    std::vector<int> order(sched.order, sched.order + M), step_ptr(sched.step_ptr, sched.step_ptr + sched.nsteps + 1);
    nbgpu_free_schedule(&sched);
    std::vector<int> row_ptr(code->row_ptr, code->row_ptr + M + 1), col(code->col, code->col + E);
    if (isolated.empty()) isolated.push_back(0), k.n_isolated = 0; else k.n_isolated = (int)isolated.size();

    int rc;
    if ((rc = upload(c, &c->d_row_ptr, row_ptr)) || (rc = upload(c, &c->d_col, col)) || (rc = upload(c, &c->d_order, order)) ||
        (rc = upload(c, &c->d_step_ptr, step_ptr)) || (rc = upload(c, &c->d_isolated, isolated)) || (rc = upload(c, &c->d_hval, hval)) ||
        (rc = upload(c, &c->d_last, last)) || (rc = upload(c, &c->d_rotin, rotin)) || (rc = upload(c, &c->d_rotout, rotout)) ||
        (rc = upload(c, &c->d_img, img)) || (rc = upload(c, &c->d_inv, inv))) { nbgpu_destroy(c); return rc; }
    k.row_ptr = c->d_row_ptr; k.col = c->d_col; k.order = c->d_order; k.step_ptr = c->d_step_ptr; k.isolated = c->d_isolated;
    k.hval = c->d_hval; k.last = c->d_last; k.rotin = c->d_rotin; k.rotout = c->d_rotout; k.img = c->d_img; k.inv = c->d_inv;

    /* occupancy -> persistent grid */
    const void *fn = q == 16 ? (const void *)decode_kernel<16> : q == 64 ? (const void *)decode_kernel<64> : (const void *)decode_kernel<256>;
    CK(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem_bytes));
    const void *fn2 = q == 16 ? (const void *)checknode_kernel<16> : q == 64 ? (const void *)checknode_kernel<64> : (const void *)checknode_kernel<256>;
    CK(c, cudaFuncSetAttribute(fn2, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem_bytes));
    int per_sm = 0;
    if (q == 16) { CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_kernel<16>, NT, k.smem_bytes)); }
    else if (q == 64) { CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_kernel<64>, NT, k.smem_bytes)); }
    else { CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_kernel<256>, NT, k.smem_bytes)); }
    if (per_sm < 1) { ctx_err(c, "decode kernel does not fit on an SM (smem %d bytes)", k.smem_bytes); nbgpu_destroy(c); return NBGPU_ECUDA; }
    if (getenv("NBGPU_CTAS_PER_SM")) per_sm = std::min(per_sm, atoi(getenv("NBGPU_CTAS_PER_SM")));
    c->per_sm = per_sm;
    c->grid = prop.multiProcessorCount * per_sm;
    const int groups = (max_batch + k.F - 1) / k.F;
    if (c->grid > groups) c->grid = groups;
    c->nslots = c->grid * k.F;

    CK(c, cudaMalloc((void **)&c->d_app, (size_t)c->nslots * N * q * sizeof(float)));
    CK(c, cudaMalloc((void **)&c->d_ctov, (size_t)c->nslots * E * k.rec_stride));
    CK(c, cudaMalloc((void **)&c->d_dec, (size_t)c->nslots * N));
    CK(c, cudaMemset(c->d_dec, 0, (size_t)c->nslots * N));
    CK(c, cudaMalloc((void **)&c->d_decide, (size_t)max_batch * N * sizeof(int)));
    CK(c, cudaMalloc((void **)&c->d_synd, (size_t)max_batch * sizeof(int)));
    CK(c, cudaMalloc((void **)&c->d_iters, (size_t)max_batch * sizeof(int)));
    CK(c, cudaMalloc((void **)&c->d_frame_slot, (size_t)max_batch * sizeof(int)));
    CK(c, cudaMalloc((void **)&c->d_slot_frame, (size_t)c->nslots * sizeof(int)));
    CK(c, cudaMemset(c->d_slot_frame, 0xff, (size_t)c->nslots * sizeof(int)));
    CK(c, cudaMalloc((void **)&c->d_queue, 2 * sizeof(unsigned)));
    c->d_slow = c->d_queue + 1;
    CK(c, cudaMemset(c->d_queue, 0, 2 * sizeof(unsigned)));
    k.app = c->d_app; k.ctov = c->d_ctov; k.dec = c->d_dec;
    k.out_decide = c->d_decide; k.out_synd = c->d_synd; k.out_iters = c->d_iters;
    k.frame_slot = c->d_frame_slot; k.slot_frame = c->d_slot_frame; k.queue = c->d_queue; k.slow_counter = c->d_slow;
    *out = c;
    return NBGPU_OK;
}

extern "C" void nbgpu_destroy(nbgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    void *bufs[] = { c->d_row_ptr, c->d_col, c->d_order, c->d_step_ptr, c->d_isolated, c->d_hval, c->d_last, c->d_rotin,
                     c->d_rotout, c->d_img, c->d_inv, c->d_app, c->d_ctov, c->d_dec, c->d_in, c->d_decide, c->d_synd,
                     c->d_iters, c->d_frame_slot, c->d_slot_frame, c->d_queue };
    for (void *b : bufs) if (b) cudaFree(b);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->stream) cudaStreamDestroy(c->stream);
    free(c->row_ptr_h); free(c->inv_h);
    free(c);
}

static int ensure_input(nbgpu_ctx *c, size_t floats)
{
    if (floats <= c->in_capacity) return NBGPU_OK;
    if (c->d_in) cudaFree(c->d_in);
    c->d_in = NULL; c->in_capacity = 0;
    CK(c, cudaMalloc((void **)&c->d_in, floats * sizeof(float)));
    c->in_capacity = floats;
    return NBGPU_OK;
}

static int upload_common(nbgpu_ctx *c, const float *src, size_t per_frame, int B, int kind)
{
    if (!c || !src) { ctx_err(c, "NULL argument"); return NBGPU_EINVAL; }
    if (B < 1 || B > c->max_batch) { ctx_err(c, "B=%d outside 1..max_batch=%d", B, c->max_batch); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    int rc = ensure_input(c, per_frame * (size_t)c->max_batch);
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(c->d_in, src, per_frame * B * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->resident_B = B; c->resident_kind = kind;
    return NBGPU_OK;
}

extern "C" int nbgpu_upload_noisy(nbgpu_ctx *c, const float *noisy, float sigma, int B)
{
    if (!c) return NBGPU_EINVAL;
    c->k.den = 2.0 * (double)(float)(sigma * sigma);                      /* 2.0*SQR(sigma), channel.c:73 */
    return upload_common(c, noisy, (size_t)c->N * c->logq, B, 0);
}
extern "C" int nbgpu_upload_llr(nbgpu_ctx *c, const float *llr, int B)
{
    if (!c) return NBGPU_EINVAL;
    return upload_common(c, llr, (size_t)c->N * c->q, B, 1);
}

extern "C" int nbgpu_run(nbgpu_ctx *c)
{
    if (!c || c->resident_B < 1) { ctx_err(c, "nbgpu_run: no resident batch (call nbgpu_upload_* first)"); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    KArgs k = c->k;
    k.B = c->resident_B; k.input_kind = c->resident_kind; k.in = c->d_in;
    CK(c, cudaMemsetAsync(c->d_queue, 0, sizeof(unsigned), c->stream));
    const int groups = (k.B + k.F - 1) / k.F;
    const int grid = std::min(c->grid, groups);
    CK(c, cudaEventRecord(c->ev0, c->stream));
    if (c->q == 16) decode_kernel<16><<<grid, NT, k.smem_bytes, c->stream>>>(k);
    else if (c->q == 64) decode_kernel<64><<<grid, NT, k.smem_bytes, c->stream>>>(k);
    else decode_kernel<256><<<grid, NT, k.smem_bytes, c->stream>>>(k);
    CK(c, cudaGetLastError());
    CK(c, cudaEventRecord(c->ev1, c->stream));
    c->launches += 1;
    return NBGPU_OK;
}

extern "C" int nbgpu_sync(nbgpu_ctx *c)
{
    if (!c) return NBGPU_EINVAL;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    return NBGPU_OK;
}

extern "C" int nbgpu_last_kernel_ms(nbgpu_ctx *c, float *ms)
{
    if (!c || !ms) return NBGPU_EINVAL;
    CK(c, cudaEventSynchronize(c->ev1));
    CK(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return NBGPU_OK;
}
extern "C" long nbgpu_launch_count(const nbgpu_ctx *c) { return c ? c->launches : 0; }

extern "C" int nbgpu_timer_begin(nbgpu_ctx *c)
{
    if (!c) return NBGPU_EINVAL;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaEventRecord(c->ev_t0, c->stream));
    return NBGPU_OK;
}
extern "C" int nbgpu_timer_end(nbgpu_ctx *c, float *ms)
{
    if (!c || !ms) return NBGPU_EINVAL;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaEventRecord(c->ev_t1, c->stream));
    CK(c, cudaEventSynchronize(c->ev_t1));
    CK(c, cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return NBGPU_OK;
}
extern "C" int nbgpu_host_register(void *ptr, size_t bytes)
{
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); ctx_err(NULL, "cudaHostRegister failed: %s", cudaGetErrorString(e)); return NBGPU_ECUDA; }
    return NBGPU_OK;
}
extern "C" int nbgpu_host_unregister(void *ptr)
{
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); ctx_err(NULL, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); return NBGPU_ECUDA; }
    return NBGPU_OK;
}
extern "C" int nbgpu_geometry(const nbgpu_ctx *c, int *geo)
{
    if (!c || !geo) return NBGPU_EINVAL;
    geo[0] = c->grid; geo[1] = c->k.F; geo[2] = c->k.G / (c->k.F > 0 ? c->k.F : 1); geo[3] = c->k.nsteps;
    geo[4] = c->k.smem_bytes; geo[5] = c->nslots; geo[6] = c->per_sm; geo[7] = c->k.rec_stride;
    return NBGPU_OK;
}
extern "C" long nbgpu_slow_selects(nbgpu_ctx *c)
{
    if (!c) return -1;
    unsigned v = 0;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess ||
        cudaMemcpy(&v, c->d_slow, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (long)v;
}

extern "C" int nbgpu_download(nbgpu_ctx *c, int *decide, int *synd, int *iters)
{
    if (!c || c->resident_B < 1) { ctx_err(c, "nbgpu_download: nothing decoded"); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    const int B = c->resident_B;
    if (decide) CK(c, cudaMemcpyAsync(decide, c->d_decide, (size_t)B * c->N * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (synd) CK(c, cudaMemcpyAsync(synd, c->d_synd, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (iters) CK(c, cudaMemcpyAsync(iters, c->d_iters, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return NBGPU_OK;
}

extern "C" int nbgpu_decode_noisy(nbgpu_ctx *c, const float *noisy, float sigma, int B, int *decide, int *synd, int *iters)
{
    int rc;
    if ((rc = nbgpu_upload_noisy(c, noisy, sigma, B)) || (rc = nbgpu_run(c)) || (rc = nbgpu_download(c, decide, synd, iters))) return rc;
    return NBGPU_OK;
}
extern "C" int nbgpu_decode_llr(nbgpu_ctx *c, const float *llr, int B, int *decide, int *synd, int *iters)
{
    int rc;
    if ((rc = nbgpu_upload_llr(c, llr, B)) || (rc = nbgpu_run(c)) || (rc = nbgpu_download(c, decide, synd, iters))) return rc;
    return NBGPU_OK;
}

extern "C" int nbgpu_get_state(nbgpu_ctx *c, int frame, float *APP, float *CtoV)
{
    if (!c || frame < 0 || frame >= c->resident_B) { ctx_err(c, "nbgpu_get_state: frame %d not in the last batch", frame); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    int slot = -1, owner = -1;
    CK(c, cudaMemcpy(&slot, c->d_frame_slot + frame, sizeof(int), cudaMemcpyDeviceToHost));
    if (slot < 0 || slot >= c->nslots) { ctx_err(c, "frame %d has no slot", frame); return NBGPU_ESTATE; }
    CK(c, cudaMemcpy(&owner, c->d_slot_frame + slot, sizeof(int), cudaMemcpyDeviceToHost));
    if (owner != frame) { ctx_err(c, "state of frame %d was overwritten by frame %d (batch larger than the resident slots)", frame, owner); return NBGPU_ESTATE; }
    const int N = c->N, q = c->q, E = c->E, n_m = c->p.n_m, rs = c->k.rec_stride;
    if (APP) CK(c, cudaMemcpy(APP, c->d_app + (size_t)slot * N * q, (size_t)N * q * sizeof(float), cudaMemcpyDeviceToHost));
    if (CtoV) {
        std::vector<uint8_t> rec((size_t)E * rs);
        CK(c, cudaMemcpy(rec.data(), c->d_ctov + (size_t)slot * E * rs, rec.size(), cudaMemcpyDeviceToHost));
        for (int e = 0; e < E; e++) {
            const uint8_t *r = rec.data() + (size_t)e * rs;
            float sat; int stp;
            memcpy(&sat, r + 4 * n_m, 4); memcpy(&stp, r + 4 * n_m + 4, 4);
            for (int g = 0; g < q; g++) CtoV[(size_t)e * q + g] = sat;
            for (int kk = 0; kk < stp; kk++) { float l; memcpy(&l, r + 4 * kk, 4); CtoV[(size_t)e * q + r[4 * n_m + 8 + kk]] = l; }
        }
    }
    return NBGPU_OK;
}

/* ---- unit boundaries ---- */
template <typename T> struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)); }
};

extern "C" int nbgpu_select_nm(nbgpu_ctx *c, const float *rows, float *llr, int *gf, int B)
{
    if (!c || !rows || !llr || !gf || B < 1) { ctx_err(c, "nbgpu_select_nm: bad argument"); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    const int q = c->q, n_m = c->p.n_m;
    DevBuf<float> d_rows, d_llr; DevBuf<int> d_gf;
    CK(c, d_rows.alloc((size_t)B * q)); CK(c, d_llr.alloc((size_t)B * n_m)); CK(c, d_gf.alloc((size_t)B * n_m));
    CK(c, cudaMemcpyAsync(d_rows.p, rows, (size_t)B * q * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const int grid = std::min((B + NW - 1) / NW, 148 * 8);
    if (q == 16) select_kernel<16><<<grid, NT, 0, c->stream>>>(d_rows.p, d_llr.p, d_gf.p, B, n_m, c->d_slow);
    else if (q == 64) select_kernel<64><<<grid, NT, 0, c->stream>>>(d_rows.p, d_llr.p, d_gf.p, B, n_m, c->d_slow);
    else select_kernel<256><<<grid, NT, 0, c->stream>>>(d_rows.p, d_llr.p, d_gf.p, B, n_m, c->d_slow);
    CK(c, cudaGetLastError());
    c->launches++;
    CK(c, cudaMemcpyAsync(llr, d_llr.p, (size_t)B * n_m * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(gf, d_gf.p, (size_t)B * n_m * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return NBGPU_OK;
}

extern "C" int nbgpu_elementary_step(nbgpu_ctx *c, const float *in1, const float *in2, const int *idx1, const int *idx2,
                                     float *out, int *idxout, int B)
{
    if (!c || !in1 || !in2 || !idx1 || !idx2 || !out || !idxout || B < 1) { ctx_err(c, "nbgpu_elementary_step: bad argument"); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    const int n_m = c->p.n_m, q = c->q, mwords = (q + 31) / 32;
    /* symbols -> binary images + valid lengths (first -1 ends a list, bubble_decoder.c:478) */
    std::vector<uint8_t> s1((size_t)B * n_m), s2((size_t)B * n_m);
    std::vector<int> l1(B), l2(B);
    std::vector<uint8_t> img(q);
    CK(c, cudaMemcpy(img.data(), c->d_img, q, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; b++) {
        int a = n_m, d = n_m;
        for (int k = n_m - 1; k >= 0; k--) { if (idx1[(size_t)b * n_m + k] < 0) a = k; if (idx2[(size_t)b * n_m + k] < 0) d = k; }
        /* lists with a "hole" (-1 followed by a valid symbol) do not occur in the decoder; treat as ended */
        l1[b] = a; l2[b] = d;
        for (int k = 0; k < n_m; k++) {
            const int x = idx1[(size_t)b * n_m + k], y = idx2[(size_t)b * n_m + k];
            s1[(size_t)b * n_m + k] = (x >= 0 && x < q) ? img[x] : 0;
            s2[(size_t)b * n_m + k] = (y >= 0 && y < q) ? img[y] : 0;
        }
    }
    DevBuf<float> d1, d2, dout; DevBuf<uint8_t> ds1, ds2, dso; DevBuf<int> dl1, dl2, dlo; DevBuf<uint32_t> dmask;
    CK(c, d1.alloc((size_t)B * n_m)); CK(c, d2.alloc((size_t)B * n_m)); CK(c, dout.alloc((size_t)B * n_m));
    CK(c, ds1.alloc((size_t)B * n_m)); CK(c, ds2.alloc((size_t)B * n_m)); CK(c, dso.alloc((size_t)B * n_m));
    CK(c, dl1.alloc(B)); CK(c, dl2.alloc(B)); CK(c, dlo.alloc(B)); CK(c, dmask.alloc((size_t)B * mwords));
    CK(c, cudaMemcpy(d1.p, in1, (size_t)B * n_m * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(d2.p, in2, (size_t)B * n_m * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(ds1.p, s1.data(), (size_t)B * n_m, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(ds2.p, s2.data(), (size_t)B * n_m, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(dl1.p, l1.data(), (size_t)B * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(dl2.p, l2.data(), (size_t)B * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemset(dso.p, 0, (size_t)B * n_m));
    es_kernel<<<(B + 127) / 128, 128, 0, c->stream>>>(d1.p, d2.p, ds1.p, ds2.p, dl1.p, dl2.p, dout.p, dso.p, dlo.p, dmask.p, B, n_m,
                                                     c->p.nb_oper, mwords);
    CK(c, cudaGetLastError());
    c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));
    std::vector<uint8_t> so((size_t)B * n_m); std::vector<int> lo(B);
    CK(c, cudaMemcpy(out, dout.p, (size_t)B * n_m * 4, cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(so.data(), dso.p, (size_t)B * n_m, cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(lo.data(), dlo.p, (size_t)B * 4, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; b++) for (int k = 0; k < n_m; k++)
        idxout[(size_t)b * n_m + k] = (k < lo[b]) ? c->inv_h[so[(size_t)b * n_m + k]] : -1;
    return NBGPU_OK;
}

extern "C" int nbgpu_check_node(nbgpu_ctx *c, int node, const float *vllr, const int *vgf, float *cllr, int *cgf, int B)
{
    if (!c || !vllr || !vgf || !cllr || !cgf || B < 1 || node < 0 || node >= c->M) { ctx_err(c, "nbgpu_check_node: bad argument"); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    const int dc = c->row_ptr_h[node + 1] - c->row_ptr_h[node], n_m = c->p.n_m, q = c->q;
    DevBuf<float> dvl, dcl; DevBuf<int> dvg, dcg;
    CK(c, dvl.alloc((size_t)B * dc * n_m)); CK(c, dvg.alloc((size_t)B * dc * n_m));
    CK(c, dcl.alloc((size_t)B * dc * q)); CK(c, dcg.alloc((size_t)B * dc * q));
    CK(c, cudaMemcpy(dvl.p, vllr, (size_t)B * dc * n_m * 4, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(dvg.p, vgf, (size_t)B * dc * n_m * 4, cudaMemcpyHostToDevice));
    const int grid = std::min((B + c->k.G - 1) / c->k.G, 148 * 4);
    if (q == 16) checknode_kernel<16><<<grid, NT, c->k.smem_bytes, c->stream>>>(c->k, node, dvl.p, dvg.p, dcl.p, dcg.p, B);
    else if (q == 64) checknode_kernel<64><<<grid, NT, c->k.smem_bytes, c->stream>>>(c->k, node, dvl.p, dvg.p, dcl.p, dcg.p, B);
    else checknode_kernel<256><<<grid, NT, c->k.smem_bytes, c->stream>>>(c->k, node, dvl.p, dvg.p, dcl.p, dcg.p, B);
    CK(c, cudaGetLastError());
    c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(cllr, dcl.p, (size_t)B * dc * q * 4, cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(cgf, dcg.p, (size_t)B * dc * q * 4, cudaMemcpyDeviceToHost));
    return NBGPU_OK;
}

extern "C" int nbgpu_decision_syndrome(nbgpu_ctx *c, const float *app, int *decide, int *synd, int B)
{
    if (!c || !app || !decide || !synd || B < 1) { ctx_err(c, "nbgpu_decision_syndrome: bad argument"); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    const int N = c->N, q = c->q;
    DevBuf<float> dapp; DevBuf<int> ddec, dsyn;
    CK(c, dapp.alloc((size_t)B * N * q)); CK(c, ddec.alloc((size_t)B * N)); CK(c, dsyn.alloc(B));
    CK(c, cudaMemcpy(dapp.p, app, (size_t)B * N * q * 4, cudaMemcpyHostToDevice));
    const int grid = (int)std::min<long>(((long)B * N + NW - 1) / NW, 148 * 8);
    if (q == 16) decision_kernel<16><<<grid, NT, 0, c->stream>>>(c->k, dapp.p, ddec.p, B);
    else if (q == 64) decision_kernel<64><<<grid, NT, 0, c->stream>>>(c->k, dapp.p, ddec.p, B);
    else decision_kernel<256><<<grid, NT, 0, c->stream>>>(c->k, dapp.p, ddec.p, B);
    CK(c, cudaGetLastError());
    syndrome_kernel<<<std::min(B, 148 * 4), NT, 0, c->stream>>>(c->k, ddec.p, dsyn.p, B);
    CK(c, cudaGetLastError());
    c->launches += 2;
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(decide, ddec.p, (size_t)B * N * 4, cudaMemcpyDeviceToHost));
    CK(c, cudaMemcpy(synd, dsyn.p, (size_t)B * 4, cudaMemcpyDeviceToHost));
    return NBGPU_OK;
}

extern "C" int nbgpu_channel_awgn_bpsk(nbgpu_ctx *c, const float *noisy, float sigma, int B, float *llr, float *illr, int *igf)
{
    if (!c || !noisy || B < 1 || ((illr == NULL) != (igf == NULL))) { ctx_err(c, "nbgpu_channel_awgn_bpsk: bad argument"); return NBGPU_EINVAL; }
    CK(c, cudaSetDevice(c->device));
    const int N = c->N, q = c->q;
    DevBuf<float> dn, dl, dil; DevBuf<int> dig;
    CK(c, dn.alloc((size_t)B * N * c->logq)); CK(c, dl.alloc((size_t)B * N * q));
    if (illr) { CK(c, dil.alloc((size_t)B * N * q)); CK(c, dig.alloc((size_t)B * N * q)); }
    CK(c, cudaMemcpy(dn.p, noisy, (size_t)B * N * c->logq * 4, cudaMemcpyHostToDevice));
    KArgs k = c->k;
    k.den = 2.0 * (double)(float)(sigma * sigma);
    const int grid = (int)std::min<long>(((long)B * N + NW - 1) / NW, 148 * 8);
    if (q == 16) channel_kernel<16><<<grid, NT, 0, c->stream>>>(k, dn.p, dl.p, illr ? dil.p : nullptr, illr ? dig.p : nullptr, B);
    else if (q == 64) channel_kernel<64><<<grid, NT, 0, c->stream>>>(k, dn.p, dl.p, illr ? dil.p : nullptr, illr ? dig.p : nullptr, B);
    else channel_kernel<256><<<grid, NT, 0, c->stream>>>(k, dn.p, dl.p, illr ? dil.p : nullptr, illr ? dig.p : nullptr, B);
    CK(c, cudaGetLastError());
    c->launches++;
    CK(c, cudaStreamSynchronize(c->stream));
    if (llr) CK(c, cudaMemcpy(llr, dl.p, (size_t)B * N * q * 4, cudaMemcpyDeviceToHost));
    if (illr) {
        CK(c, cudaMemcpy(illr, dil.p, (size_t)B * N * q * 4, cudaMemcpyDeviceToHost));
        CK(c, cudaMemcpy(igf, dig.p, (size_t)B * N * q * 4, cudaMemcpyDeviceToHost));
    }
    return NBGPU_OK;
}
