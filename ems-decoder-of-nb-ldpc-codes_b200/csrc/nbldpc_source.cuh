/* nbldpc_source.cuh -- the reference's Monte-Carlo frame source on the device.
 *
 * Replaces, for a whole batch of frames and without touching the host,
 *   RandomBinaryGenerator   tools.c:124-136   information bits  floor((float)drand48() * 1.9999)
 *   Encoding                tools.c:232-270   back-substitution on the triangular matrix of GaussianElimination
 *                                             (tools.c:151-218; done once on the host, nbldpc_host.c)
 *   ModelChannel_AWGN_BPSK  channel.c:51-62   y = BPSK(b) + sigma*sqrt(-2 ln u)*cos(2 pi v), f64 -> f32
 *   the error count         NB_LDPC.c:479-485 information-bit errors of the decided word
 *
 * Frame f of the stream owns the drand48 draws [f*D, (f+1)*D), D = (K + 2N) log2 q, so every thread reaches its own
 * draws by LCG jump-ahead (48 precomputed powers of the affine map in constant memory).
 *
 * Exactness.  Everything except log() and cos() is IEEE arithmetic that the device reproduces bit for bit (explicit
 * __dmul_rn/__dadd_rn so that nothing is contracted into an FMA).  The device's log/cos may differ from glibc's by a
 * few ulp of a double, which changes the f32 sample only when the f64 value sits within that distance of an f32
 * rounding boundary.  Such samples (about one in 10^6) are FLAGGED by a margin test that is ~8x wider than the
 * documented error bounds, and recomputed by the host with glibc (nbgpu_noise_sample, nbldpc_host.c) before the
 * decoder reads them.  tests/test_gpu_source.py checks the stream against the host source sample by sample.
 *
 * GF arithmetic of the encoder runs on binary images (addition = XOR): mulimg[h][x] = img(h * sym(x)),
 * divimg[p][x] = img(sym(x) / p), both built on the host from the code's tables.
 */
#pragma once
#include <stdint.h>

#define SRC_MASK ((1ULL << 48) - 1)
#define SRC_A 0x5DEECE66DULL
#define SRC_C 0xBULL

__constant__ uint64_t c_src_mul[48];     /* a^(2^j) mod 2^48 */
__constant__ uint64_t c_src_add[48];     /* increment of the 2^j-fold map */

struct SrcArgs {
    int N, M, K, q, logq, B;
    uint64_t x0;                 /* generator state before the first draw of this batch's first frame */
    uint64_t D;                  /* draws per frame */
    int nlevels;
    const int *level_ptr;        /* [nlevels+1] into row_order: rows of one level only need rows of earlier levels */
    const int *row_order;        /* [M] */
    const int *ut_ptr, *ut_col;  /* sparse rows of the triangular matrix (strictly right of the diagonal) */
    const uint8_t *ut_val, *piv; /* coefficients (symbols), diagonal */
    const int *perm;             /* [N] column permutation of the elimination */
    const uint8_t *mulimg, *divimg;
    const uint8_t *img;          /* [q] binary image of a symbol */
    uint8_t *cw;                 /* [B][N] codeword, binary images */
    float *noisy;                /* [B][N*logq] */
    float sigma;
    double margin;
    unsigned *flag_count, *flags;
    unsigned flag_cap;
    const int *decide;           /* [B][N] */
    int *bit_errors;             /* [B] */
};

__device__ __forceinline__ uint64_t src_skip(uint64_t x, uint64_t n)
{
    for (int j = 0; n; ++j, n >>= 1)
        if (n & 1) x = (x * c_src_mul[j] + c_src_add[j]) & SRC_MASK;
    return x;
}
/* (float)drand48(), tools.c:73 */
__device__ __forceinline__ float src_draw(uint64_t &x)
{
    x = (x * SRC_A + SRC_C) & SRC_MASK;
    return __double2float_rn(__dmul_rn((double)(long long)x, 1.0 / 281474976710656.0));
}

/* One CTA per frame: information symbols, then the back-substitution level by level (a warp per row). */
__global__ void __launch_bounds__(256) source_encode_kernel(SrcArgs a)
{
    extern __shared__ uint8_t ns[];                       /* [N] symbols in elimination order, binary images */
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint64_t xf = src_skip(a.x0, (uint64_t)f * a.D);
    for (int k = tid; k < a.K; k += blockDim.x) {
        uint64_t x = src_skip(xf, (uint64_t)k * a.logq);
        int bits = 0;
        for (int l = 0; l < a.logq; l++) {
            const float u = src_draw(x);
            bits |= (__dmul_rn((double)u, 1.9999) >= 1.0 ? 1 : 0) << l;       /* floor(u*1.9999), tools.c:132 */
        }
        ns[a.M + k] = (uint8_t)bits;                                          /* Bin2GF then BINGF: the image is the bits */
    }
    __syncthreads();
    for (int lv = 0; lv < a.nlevels; lv++) {
        for (int r = a.level_ptr[lv] + warp; r < a.level_ptr[lv + 1]; r += nwarps) {
            const int m = a.row_order[r];
            unsigned acc = 0;
            for (int k = a.ut_ptr[m] + lane; k < a.ut_ptr[m + 1]; k += 32)
                acc ^= a.mulimg[(int)a.ut_val[k] * a.q + ns[a.ut_col[k]]];     /* tools.c:249-250 */
            acc = __reduce_xor_sync(0xffffffffu, acc);
            if (lane == 0) ns[m] = a.divimg[(int)a.piv[m] * a.q + acc];         /* tools.c:253 */
        }
        __syncthreads();
    }
    uint8_t *cw = a.cw + (size_t)f * a.N;
    for (int n = tid; n < a.N; n += blockDim.x) cw[a.perm[n]] = ns[n];          /* tools.c:257-258 */
}

/* One thread per codeword symbol: its 2*logq draws, log2 q noisy samples. */
__global__ void __launch_bounds__(256) source_noise_kernel(SrcArgs a)
{
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)a.B * a.N) return;
    const int f = (int)(id / a.N), n = (int)(id % a.N);
    uint64_t x = src_skip(a.x0, (uint64_t)f * a.D + (uint64_t)a.K * a.logq + (uint64_t)n * 2 * a.logq);
    const int sym = a.cw[id];
    const double sigma = (double)a.sigma;
    float *out = a.noisy + (size_t)id * a.logq;
    for (int l = 0; l < a.logq; l++) {
        const float u = src_draw(x);
        const float v = src_draw(x);
        const int bit = (sym >> l) & 1;
        const double r = sqrt(-2.0 * log((double)u));
        const double c = cos(__dmul_rn(2.0 * 3.1415926536, (double)v));          /* PI of channel.c:18 */
        const double t = __dmul_rn(__dmul_rn(sigma, r), c);
        const double y = __dadd_rn((double)(1 - 2 * bit), t);                    /* channel.c:59 */
        const float yf = __double2float_rn(y);
        const double d = a.margin * (fabs(t) + fabs(y));
        if (!(__double2float_rn(y + d) == __double2float_rn(y - d))) {            /* also true for inf/NaN */
            const unsigned slot = atomicAdd(a.flag_count, 1u);
            if (slot < a.flag_cap) a.flags[slot] = (unsigned)(id * a.logq + l) | ((unsigned)bit << 31);
        }
        out[l] = yf;
    }
}

__global__ void source_patch_kernel(float *noisy, const unsigned *idx, const float *val, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) noisy[idx[i] & 0x7fffffffu] = val[i];
}

/* One CTA per frame: information-bit errors of the decided word (NB_LDPC.c:479-485: the first K symbols). */
__global__ void __launch_bounds__(256) source_errors_kernel(SrcArgs a)
{
    __shared__ int part[8];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int *dec = a.decide + (size_t)f * a.N;
    const uint8_t *cw = a.cw + (size_t)f * a.N;
    int e = 0;
    for (int k = tid; k < a.K; k += blockDim.x) e += __popc((unsigned)a.img[dec[k]] ^ (unsigned)cw[k]);
    e = __reduce_add_sync(0xffffffffu, e);
    if ((tid & 31) == 0) part[tid >> 5] = e;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += part[w];
        a.bit_errors[f] = s;
    }
}
