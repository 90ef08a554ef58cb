/*
 * nbldpc_device.cuh -- device functions of the EMS decoder (sm_100a).
 *
 * Building blocks (reference lines they replace):
 *   warp_select_nm   NB_LDPC.c:354-374   stable top-n_m of q values + normalisation, one warp per edge
 *   es_serial        bubble_decoder.c:316-593  ElementaryStep, one THREAD per step (32 steps per warp)
 *   warp_argmin      tools.c:312-330     Decision
 * GF symbols travel through the check node as BINARY IMAGES so that GF addition is XOR
 * (ADDGF[a][b] == inv[img[a]^img[b]] is verified on the host); the multiplication by the edge
 * coefficient on the way in (bubble_decoder.c:133-152) and the division on the way out (:249-254) are
 * fused with the image mapping into two q x q byte tables (rotin / rotout).
 *
 * All LLR arithmetic uses the explicit round-to-nearest intrinsics (__fadd_rn/__fsub_rn): one IEEE
 * f32 operation where the reference has one, never contracted or re-associated.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NB_SENT 1e5f
#define NB_FULL 0xffffffffu
#define NB_KEY_INF 0xffffffffu

template <int Q> struct QTraits {
    static constexpr int VPL = (Q >= 32) ? Q / 32 : 1;      /* values per lane when a warp holds one row */
    static constexpr int LOGQ = (Q == 16) ? 4 : (Q == 64) ? 6 : 8;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

/* ---- row I/O: a warp moves one q-float row; lane holds symbols lane*VPL .. lane*VPL+VPL-1 ---- */
template <int Q> __device__ __forceinline__ void load_row(const float *row, int lane, float (&v)[QTraits<Q>::VPL])
{
    if constexpr (Q == 256) {
        const float4 a = reinterpret_cast<const float4 *>(row)[lane * 2];
        const float4 b = reinterpret_cast<const float4 *>(row)[lane * 2 + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (Q == 64) {
        const float2 a = reinterpret_cast<const float2 *>(row)[lane];
        v[0] = a.x; v[1] = a.y;
    } else {
        v[0] = (lane < Q) ? row[lane] : NB_SENT;
    }
}
template <int Q> __device__ __forceinline__ void store_row(float *row, int lane, const float (&v)[QTraits<Q>::VPL])
{
    if constexpr (Q == 256) {
        reinterpret_cast<float4 *>(row)[lane * 2] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(row)[lane * 2 + 1] = make_float4(v[4], v[5], v[6], v[7]);
    } else if constexpr (Q == 64) {
        reinterpret_cast<float2 *>(row)[lane] = make_float2(v[0], v[1]);
    } else {
        if (lane < Q) row[lane] = v[0];
    }
}

/* ---- in-lane sorting networks on u32 keys ---- */
__device__ __forceinline__ void cswap(uint32_t &a, uint32_t &b)
{
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
template <int VPL> __device__ __forceinline__ void sort_keys(uint32_t (&k)[VPL])
{
    if constexpr (VPL == 8) {
        cswap(k[0], k[1]); cswap(k[2], k[3]); cswap(k[4], k[5]); cswap(k[6], k[7]);
        cswap(k[0], k[2]); cswap(k[1], k[3]); cswap(k[4], k[6]); cswap(k[5], k[7]);
        cswap(k[1], k[2]); cswap(k[5], k[6]); cswap(k[0], k[4]); cswap(k[3], k[7]);
        cswap(k[1], k[5]); cswap(k[2], k[6]);
        cswap(k[1], k[4]); cswap(k[3], k[6]);
        cswap(k[2], k[4]); cswap(k[3], k[5]);
        cswap(k[3], k[4]);
    } else if constexpr (VPL == 2) {
        cswap(k[0], k[1]);
    }
}

/* per-warp scratch in shared memory */
template <int Q> struct WarpScratch {
    uint32_t sorted[(QTraits<Q>::VPL + 1) * 32];   /* lane-sorted keys, [r][lane]; row VPL = +inf      */
    float row[Q < 64 ? 64 : Q];                    /* one dense q-row (also 2*logq doubles at intake) */
    uint32_t sel[36];                              /* winners of the selection rounds                  */
};

/* lexicographic (value, symbol) minimum across the warp with the reference's scan semantics:
 * strict '<' from the +1e5 sentinel, ties -> lowest symbol; "nothing below 1e5" -> (1e5, BIG). */
__device__ __forceinline__ void warp_lexmin(float &bv, int &bg)
{
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const float ov = __shfl_xor_sync(NB_FULL, bv, off);
        const int og = __shfl_xor_sync(NB_FULL, bg, off);
        if (ov < bv || (ov == bv && og < bg)) { bv = ov; bg = og; }
    }
}

/* Decision for one row held by a warp (tools.c:317-329) */
template <int Q> __device__ __forceinline__ int warp_argmin(const float (&v)[QTraits<Q>::VPL], int lane)
{
    constexpr int VPL = QTraits<Q>::VPL;
    float bv = NB_SENT; int bg = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < VPL; j++) if ((Q >= 32 || lane < Q) && v[j] < bv) { bv = v[j]; bg = lane * VPL + j; }
    warp_lexmin(bv, bg);
    return (bg == 0x7fffffff) ? 0 : bg;
}

/*
 * Truncation of one V->C message (NB_LDPC.c:354-374): the n_m smallest of the q values mvc[], in
 * ascending order, ties -> lowest symbol, values >= 1e5 never selected (slot keeps (1e5, symbol 0) and
 * symbol 0 is masked), then LLR[k] -= LLR[0], LLR[0] = 0.
 *
 * Fast path: every value is turned into a UNIQUE 32-bit key  (f32 bits & ~(q-1)) | symbol.  For
 * non-negative finite floats the bit pattern is monotone in the value, so key order == (value, symbol)
 * order except when two values differ only in the log2(q) dropped mantissa bits.  Each lane sorts its
 * VPL keys with a register network and parks them in shared memory; n_m+1 rounds of one REDUX.MIN
 * over the lane heads then pop the global minimum, the owning lane advancing its head pointer.  The
 * result is accepted only if (a) no value was negative/NaN/Inf, (b) all adjacent winners (including
 * the (n_m+1)-th) differ in their kept bits, (c) the n_m winners are < 1e5.  Otherwise the warp
 * re-runs the exact scan (slow path, same semantics as the reference loop).
 *
 * Result: lane k < n_m returns (llr, sym) = k-th entry.  mvc row must already be in ws.row.
 */
template <int Q>
__device__ __forceinline__ void warp_select_nm(const float (&mvc)[QTraits<Q>::VPL], int lane, WarpScratch<Q> &ws,
                                               int n_m, float &out_llr, int &out_sym, unsigned *slow_counter)
{
    constexpr int VPL = QTraits<Q>::VPL;
    constexpr int LOGQ = QTraits<Q>::LOGQ;
    const bool active = (Q >= 32) || lane < Q;
    uint32_t key[VPL];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < VPL; j++) {
        const uint32_t b = __float_as_uint(mvc[j]);
        bad |= active && (b >= 0x7f800000u);
        key[j] = active ? ((b & ~uint32_t(Q - 1)) | uint32_t(lane * VPL + j)) : NB_KEY_INF;
    }
    sort_keys<VPL>(key);
#pragma unroll
    for (int j = 0; j < VPL; j++) ws.sorted[j * 32 + lane] = key[j];
    ws.sorted[VPL * 32 + lane] = NB_KEY_INF;
    uint32_t head = key[0];
    int p = 0;
    for (int r = 0; r <= n_m; r++) {
        const uint32_t m = __reduce_min_sync(NB_FULL, head);
        if (head == m) {
            ws.sel[r] = m;
            p = min(p + 1, VPL);
            head = ws.sorted[p * 32 + lane];
        }
    }
    __syncwarp();
    const uint32_t mine = ws.sel[min(lane, n_m)];
    const uint32_t nxt = ws.sel[min(lane + 1, n_m)];
    const bool amb = lane < n_m && ((mine >> LOGQ) == (nxt >> LOGQ));
    int sym = int(mine & uint32_t(Q - 1));
    float val = ws.row[sym];
    const bool over = lane < n_m && !(val < NB_SENT);
    if (__any_sync(NB_FULL, bad || amb || over)) {
        /* exact scan, NB_LDPC.c:356-369 */
        if (slow_counter && lane == 0) atomicAdd(slow_counter, 1u);
        float tmp[VPL];
#pragma unroll
        for (int j = 0; j < VPL; j++) tmp[j] = mvc[j];
        for (int k = 0; k < n_m; k++) {
            float bv = NB_SENT; int bg = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < VPL; j++) if (active && tmp[j] < bv) { bv = tmp[j]; bg = lane * VPL + j; }
            warp_lexmin(bv, bg);
            if (bg == 0x7fffffff) bg = 0;                      /* nothing below 1e5: (1e5, symbol 0) */
#pragma unroll
            for (int j = 0; j < VPL; j++) if (bg == lane * VPL + j) tmp[j] = NB_SENT;   /* NB_LDPC.c:368 */
            if (lane == k) { val = bv; sym = bg; }
        }
    }
    const float v0 = __shfl_sync(NB_FULL, val, 0);
    out_llr = (lane == 0) ? 0.0f : __fsub_rn(val, v0);                                  /* NB_LDPC.c:372-373 */
    out_sym = sym;
}

/*
 * ElementaryStep (bubble_decoder.c:316-593), one thread per step.
 * Lists are (llr[n_m], sym[n_m], len): entries >= len are "absent" (reference: LLR 1e5, symbol -1);
 * their LLR slot holds 1e5 so that sums with them are >= 1e5 exactly as in tab_aux.
 * Symbols are binary images: ADDGF == XOR.  mask = q/32 words at stride mstride (shared memory).
 * Eight bubbles (nb_bubble = 8, :327): p < 4 walks row p to the right, p >= 4 walks column p-4 down
 * from row 4 (:445-460, :548-555).  Pop = first strictly smallest candidate below 1e5, else bubble 0
 * (minimum(), :38-56).
 */
__device__ __forceinline__ int es_serial(const float *__restrict__ l1, const uint8_t *__restrict__ s1, int len1,
                                         const float *__restrict__ l2, const uint8_t *__restrict__ s2, int len2,
                                         float *__restrict__ lo, uint8_t *__restrict__ so,
                                         uint32_t *mask, int mstride, int mwords, int n_m, int nb_oper)
{
    float bv[8];
    for (int w = 0; w < mwords; w++) mask[w * mstride] = 0u;
    const float a0 = l1[0], a1 = l1[1], a2 = l1[2], a3 = l1[3], a4 = l1[4];
    bv[0] = __fadd_rn(a0, l2[0]); bv[1] = __fadd_rn(a1, l2[0]); bv[2] = __fadd_rn(a2, l2[0]); bv[3] = __fadd_rn(a3, l2[0]);
    bv[4] = __fadd_rn(a4, l2[0]); bv[5] = __fadd_rn(a4, l2[1]); bv[6] = __fadd_rn(a4, l2[2]); bv[7] = __fadd_rn(a4, l2[3]);
    /* walk coordinate of each bubble, 8 bits each: column j for p<4 (starts 0), row i for p>=4 (starts 4) */
    unsigned long long pos = 0x0404040400000000ull;
    int s = 0;
    for (int ss = 0; ss < nb_oper; ss++) {
        float best = NB_SENT; int bp = 0;
#pragma unroll
        for (int p = 0; p < 8; p++) if (bv[p] < best) { best = bv[p]; bp = p; }
        const float val = (best < NB_SENT) ? best : bv[0];
        const int c = int((pos >> (8 * bp)) & 0xffull);
        const int i = (bp < 4) ? bp : c;
        const int j = (bp < 4) ? c : bp - 4;
        if (i >= len1 || j >= len2) break;                                  /* :478-484 */
        const int g = s1[i] ^ s2[j];                                        /* :486 */
        const uint32_t w = mask[(g >> 5) * mstride], bit = 1u << (g & 31);
        if (!(w & bit)) {                                                   /* :490-496 */
            lo[s] = val; so[s] = (uint8_t)g; mask[(g >> 5) * mstride] = w | bit; s++;
        }
        if (s == n_m) break;                                                /* :502 */
        if (i >= n_m - 1 || j >= n_m - 1) break;                            /* :506-544 */
        pos += 1ull << (8 * bp);                                            /* :548-555 */
        const int ni = (bp < 4) ? bp : c + 1;
        const int nj = (bp < 4) ? c + 1 : bp - 4;
        const float nv = __fadd_rn(l1[ni], l2[nj]);                         /* :557 */
#pragma unroll
        for (int p = 0; p < 8; p++) if (p == bp) bv[p] = nv;
    }
    for (int k = s; k < n_m; k++) lo[k] = NB_SENT;                          /* :370-374 */
    return s;
}
