/*
 * nbldpc_synd.cuh -- the syndrome-based check node on the GPU (sm_100a), one WARP per check node.
 *
 * Restates syndrome_ems (syndrome_decoder.c:26-284) with presorting_mvc (:289-496), sorting
 * (:1315-1334) and bayes (:2142-2211):
 *   1. presort: edges ascending by their 2nd-best LLR, then the first four by their 3rd-best LLR
 *      (selection with strict '<': ties keep the lower index first);
 *   2. one syndrome per configuration: LLR = f32 sum over the (presorted) edges of the list entry the
 *      configuration names, GF = sum (XOR of binary images) of the entries' symbols;
 *   3. STABLE sort of the syndromes by LLR -- the reference's insertion sort; here a stable LSD radix
 *      sort of the f32 bit patterns (all LLRs are non-negative) with the configuration index as payload;
 *   4. per edge d (presorted position): walk the sorted syndromes whose configuration does not deviate on
 *      d ("decorrelation"), add the edge's best symbol, first hit of a symbol sets its LLR, later hits
 *      go through bayes() in order; saturation at the LLR of decorrelated syndrome n_cv-1+3d, + offset.
 *      Hits of different symbols never interact, so a fifth stable radix pass groups the sorted list by symbol
 *      and every lane replays the hits of its own eight symbols -- same order inside a symbol, no conflicts.
 * Symbols are binary images (GF addition == XOR) exactly as in the bubble path.
 */
#pragma once
#include "nbldpc_device.cuh"

#define NB_SYND_MAX 1024                 /* configurations per check node the shared-memory plan supports */

/* a warp's shared memory for the syndrome check node, by shared-window address */
struct SyndMem {
    uint32_t lists;      /* dc lists: n_m f32 | n_m u8 (stride lstride) in ORIGINAL edge order        */
    uint32_t key[2];     /* [Spad] u32 sort keys (ping-pong)                                           */
    uint32_t pay[2];     /* [Spad] u16 configuration index (ping-pong)                                 */
    uint32_t gf;         /* [Spad] u8 syndrome symbol, indexed by configuration                        */
    uint32_t hist;       /* [256] u32 radix histogram; afterwards f32 output LLRs of the current edge,
                            indexed by binary image                                                      */
    uint32_t M;          /* [256] u32 first position of every symbol group in the symbol-grouped order    */
    uint32_t perm;       /* [16] i32: original edge at presorted position i; [16] = sat of the current edge */
    uint32_t cfg;        /* [S][dc] u8 configuration table (shared by the CTA)                          */
    int lstride, n_m, dc, S, Spad, n_cv;
};

/* bayes(M1 = new LLR, M2 = current LLR), syndrome_decoder.c:2142-2211.  The reference takes double arguments, keeps float
 * locals and multiplies by double constants.  Bit-exact equivalents used here:
 *   - M1 < M2 on doubles that are floats == the float comparison;
 *   - dif = (float)(M2 - M1): the double difference is kept (one DADD), then rounded;
 *   - (double)dif < 0.1 / 0.2 / 1 / 2  ==  dif < 0.1f / 0.2f / 1.0f / 2.0f  (0.1f and 0.2f are the first floats above 0.1, 0.2);
 *   - (float)(c * (double)min) for c = 0.5, 0.75, 0.9375: the double product of a float by a 1-4 bit constant is exact, so its
 *     rounding equals the float product; c = 0.825 is not representable and keeps the double multiplication. */
__device__ __forceinline__ float synd_bayes(float m1, float m2)
{
    const bool lt = m1 < m2;
    float mn = lt ? m1 : m2;
    const float hi = lt ? m2 : m1;
    const float dif = __double2float_rn(__dsub_rn((double)hi, (double)mn));
    if (dif < 0.1f) mn = __fmul_rn(0.5f, mn);
    else if (dif < 0.2f) mn = __fmul_rn(0.75f, mn);
    else if (dif < 1.0f) mn = __double2float_rn(__dmul_rn(0.825, (double)mn));
    else if (dif < 2.0f) mn = __fmul_rn(0.9375f, mn);
    return mn;
}

/* 5th radix pass: the digit is the syndrome's symbol (padding entries go to group 255 and sort last inside it) */
__device__ __forceinline__ uint32_t synd_symbol_digit(const SyndMem &sm, int src, int i)
{
    unsigned short p;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(p) : "r"(sm.pay[src] + 2 * i));
    return (int)p < sm.S ? lds_u8(sm.gf + p) : 255u;
}

/* steps 1-3: after this call key[0]/pay[0] hold the syndromes in the reference's sorted order */
__device__ __forceinline__ void synd_prepare(const SyndMem &sm, int lane)
{
    const int dc = sm.dc, n_m = sm.n_m;
    /* ---- presorting_mvc ---- */
    {
        const int i = lane < dc ? lane : 0;
        const float k1 = lds_f32(sm.lists + i * sm.lstride + 4);                /* M_VtoC_LLR[i][1], :320 */
        int rank = 0;
        for (int j = 0; j < dc; j++) {
            const float kj = __shfl_sync(NB_FULL, k1, j);
            rank += (kj < k1 || (kj == k1 && j < i)) ? 1 : 0;
        }
        if (lane < dc) sts_u32(sm.perm + 4 * rank, (uint32_t)i);
        __syncwarp();
        /* first 'border' = 4 presorted edges again, by M_VtoC_LLR[.][2], :377-469 */
        const int e = (int)lds_u32(sm.perm + 4 * (lane < 4 ? lane : 0));
        const float k2 = lds_f32(sm.lists + e * sm.lstride + 8);
        int rank2 = 0;
        for (int j = 0; j < 4; j++) {
            const float kj = __shfl_sync(NB_FULL, k2, j);
            rank2 += (kj < k2 || (kj == k2 && j < lane)) ? 1 : 0;
        }
        __syncwarp();
        if (lane < 4) sts_u32(sm.perm + 4 * rank2, (uint32_t)e);
        __syncwarp();
    }
    /* ---- syndromes, :64-77 ---- */
    if (dc == 4) {                                   /* the usual degree: list bases in registers, one table word per configuration */
        uint32_t lb[4];
#pragma unroll
        for (int j = 0; j < 4; j++) lb[j] = sm.lists + lds_u32(sm.perm + 4 * j) * sm.lstride;
        for (int i = lane; i < sm.Spad; i += 32) {
            uint32_t key = 0xffffffffu, pay = 0xffffu;
            if (i < sm.S) {
                const uint32_t cw = lds_u32(sm.cfg + 4 * i);
                float llr = 0.0f;
                uint32_t gf = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t c = (cw >> (8 * j)) & 255u;
                    llr = __fadd_rn(llr, lds_f32(lb[j] + 4 * c));
                    gf ^= lds_u8(lb[j] + 4 * n_m + c);
                }
                key = __float_as_uint(llr); pay = (uint32_t)i;
                sts_u8(sm.gf + i, gf);
            }
            sts_u32(sm.key[0] + 4 * i, key);
            asm volatile("st.shared.u16 [%0], %1;" :: "r"(sm.pay[0] + 2 * i), "h"((unsigned short)pay) : "memory");
        }
    } else {
        for (int i = lane; i < sm.Spad; i += 32) {
            uint32_t key = 0xffffffffu, gf = 0, pay = 0xffffu;
            if (i < sm.S) {
                float llr = 0.0f;
                for (int j = 0; j < dc; j++) {
                    const uint32_t list = sm.lists + lds_u32(sm.perm + 4 * j) * sm.lstride;
                    const uint32_t c = lds_u8(sm.cfg + i * dc + j);
                    llr = __fadd_rn(llr, lds_f32(list + 4 * c));
                    gf ^= lds_u8(list + 4 * n_m + c);
                }
                key = __float_as_uint(llr); pay = (uint32_t)i;
                sts_u8(sm.gf + i, gf);
            }
            sts_u32(sm.key[0] + 4 * i, key);
            asm volatile("st.shared.u16 [%0], %1;" :: "r"(sm.pay[0] + 2 * i), "h"((unsigned short)pay) : "memory");
        }
    }
    __syncwarp();
    /* ---- stable LSD radix sort, 4 passes of 8 bits over the f32 bit pattern (sorting(), :1315-1334, is a stable insertion
     * sort), then a 5th stable pass on the syndrome's symbol: buffer 0 ends up in the reference's sorted order, buffer 1
     * grouped by symbol with every group still ascending in (LLR, sorted position) ---- */
    const unsigned lt = (1u << lane) - 1u;
    for (int pass = 0; pass < 5; pass++) {
        const int src = pass & 1, dst = src ^ 1, shift = 8 * pass;
#pragma unroll
        for (int b = 0; b < 8; b++) sts_u32(sm.hist + 4 * (lane * 8 + b), 0u);
        __syncwarp();
        for (int i = lane; i < sm.Spad; i += 32) {
            uint32_t d;
            if (pass < 4) d = (lds_u32(sm.key[src] + 4 * i) >> shift) & 255u;
            else d = synd_symbol_digit(sm, src, i);
            asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(sm.hist + 4 * d) : "memory");
        }
        __syncwarp();
        /* exclusive prefix sum over the 256 bins: lane owns bins 8*lane .. 8*lane+7 */
        uint32_t h[8], tot = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) { h[b] = lds_u32(sm.hist + 4 * (lane * 8 + b)); tot += h[b]; }
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(NB_FULL, inc, o); if (lane >= o) inc += t; }
        uint32_t run = inc - tot;
#pragma unroll
        for (int b = 0; b < 8; b++) { sts_u32(sm.hist + 4 * (lane * 8 + b), run); run += h[b]; }
        if (pass == 4) {                                   /* group boundaries of the symbol-grouped order: start[256] + end */
#pragma unroll
            for (int b = 0; b < 8; b++) sts_u32(sm.M + 4 * (lane * 8 + b), lds_u32(sm.hist + 4 * (lane * 8 + b)));
        }
        __syncwarp();
        /* stable scatter: two chunks of 32 per iteration so that two MATCH are in flight */
        for (int i = lane; i < sm.Spad; i += 64) {
            uint32_t k[2], d[2]; unsigned short p[2]; unsigned peers[2]; bool ok[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int ii = i + 32 * u;
                ok[u] = ii < sm.Spad;
                k[u] = ok[u] ? lds_u32(sm.key[src] + 4 * ii) : 0xffffffffu;
                p[u] = 0xffff;
                if (ok[u]) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(p[u]) : "r"(sm.pay[src] + 2 * ii));
                d[u] = pass < 4 ? ((k[u] >> shift) & 255u) : (ok[u] ? synd_symbol_digit(sm, src, ii) : 255u);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) peers[u] = __match_any_sync(NB_FULL, d[u]);
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (i - lane + 32 * u < sm.Spad) {          /* warp-uniform */
                    const uint32_t base = lds_u32(sm.hist + 4 * d[u]);
                    const uint32_t pos = base + __popc(peers[u] & lt);
                    sts_u32(sm.key[dst] + 4 * pos, k[u]);
                    asm volatile("st.shared.u16 [%0], %1;" :: "r"(sm.pay[dst] + 2 * pos), "h"(p[u]) : "memory");
                    __syncwarp();
                    if ((peers[u] & lt) == 0) sts_u32(sm.hist + 4 * d[u], base + __popc(peers[u]));
                    __syncwarp();
                }
            }
        }
    }
}

/* Branch-free bayes(): every lane of the walk below evaluates it on every hit, so divergence would cost more than the
 * few extra instructions (same operations and roundings as synd_bayes). */
__device__ __forceinline__ float synd_bayes_sel(float m1, float m2)
{
    const float mn = fminf(m1, m2), hi = fmaxf(m1, m2);          /* M1 < M2 ? (M1, M2) : (M2, M1); equal values give the same pair */
    const float dif = __double2float_rn(__dsub_rn((double)hi, (double)mn));
    const float f = dif < 0.1f ? 0.5f : dif < 0.2f ? 0.75f : dif < 2.0f ? 0.9375f : 1.0f;
    const float a = __fmul_rn(f, mn);
    const float b = __double2float_rn(__dmul_rn(0.825, (double)mn));
    return (dif >= 0.2f && dif < 1.0f) ? b : a;
}

/* step 4, saturation levels: sat_d = LLR of decorrelated syndrome number n_cv-1+3d in the sorted order (:195), for every
 * presorted position d; stored as f32 at sm.perm + 64 + 4d.  Must run before synd_walk overwrites the sorted buffer. */
__device__ __forceinline__ void synd_sats(const SyndMem &sm, int lane)
{
    const int dc = sm.dc;
    const unsigned lt = (1u << lane) - 1u;
    for (int d = 0; d < dc; d++) {
        const int target = sm.n_cv - 1 + 3 * d;
        float sat = 0.0f;
        int cnt = 0;
        for (int i = lane; i < sm.Spad; i += 32) {
            unsigned short p;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(p) : "r"(sm.pay[0] + 2 * i));
            const bool keep = (int)p < sm.S && lds_u8(sm.cfg + (int)p * dc + d) == 0;         /* :96-98 */
            const unsigned bal = __ballot_sync(NB_FULL, keep);
            const int n = __popc(bal);
            if (cnt + n > target) {
                const bool mine = keep && cnt + __popc(bal & lt) == target;
                const unsigned who = __ballot_sync(NB_FULL, mine);
                sat = __shfl_sync(NB_FULL, __uint_as_float(lds_u32(sm.key[0] + 4 * i)), __ffs(who) - 1);
                break;
            }
            cnt += n;
        }
        if (lane == 0) sts_f32(sm.perm + 64 + 4 * d, sat);
    }
    __syncwarp();
}

/* step 4 for the presorted positions d0 .. d0+nd-1 (nd <= 4) in ONE walk: on return out_k[256] = sm.key[0] + 1024 k
 * (f32, indexed by the binary image of the rotated symbol) holds M_CtoV_LLR[d0+k][.] after saturation
 * (syndrome_decoder.c:93-209).  The reference walks the sorted syndromes and, per symbol, lets the first hit set the LLR
 * and every later hit go through bayes(); hits of different symbols do not interact, so lane L replays the syndromes of
 * the symbol groups 8L..8L+7 (one contiguous range of the symbol-grouped order, ascending inside a group) for all nd
 * edges at once and scatters the results to the symbols (group ^ best symbol of the edge). */
__device__ __forceinline__ void synd_walk(const SyndMem &sm, int d0, int nd, float offset, int lane)
{
    const int dc = sm.dc;
    uint32_t x[4]; float sat[4], hi[4], m[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int d = d0 + (k < nd ? k : 0);
        x[k] = lds_u8(sm.lists + lds_u32(sm.perm + 4 * d) * sm.lstride + 4 * sm.n_m);        /* M_VtoC_GF[d][0], :103 */
        sat[k] = lds_f32(sm.perm + 64 + 4 * d);
        hi[k] = __fadd_rn(sat[k], offset);
        m[k] = 0.0f;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; k++)                                  /* symbols without a hit keep the initial 1500.0 (:131), which the saturation (:198-209) turns into sat + offset unless sat >= 1500 */
        if (k < nd) {
            const uint32_t a = sm.key[0] + 1024 * k + lane * 32;
            const float unset = 1500.0f > sat[k] ? hi[k] : 1500.0f;
            asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n\tst.shared.v4.f32 [%0+16], {%1, %1, %1, %1};" :: "r"(a), "f"(unset) : "memory");
        }
    __syncwarp();
    const uint32_t g0 = (uint32_t)lane * 8u;
    const int lo = (int)lds_u32(sm.M + 4 * g0);
    const int hi_i = g0 + 8u >= 256u ? sm.Spad : (int)lds_u32(sm.M + 4 * (g0 + 8u));
    uint32_t cur = 0xffffffffu, have = 0u;
    for (int i = lo; i <= hi_i; i++) {                           /* one extra trip flushes the last group */
        uint32_t p = 0xffffu, g = 0xfffffffeu, cw = 0xffffffffu;
        float llr = 0.0f;
        if (i < hi_i) {
            unsigned short ps;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(ps) : "r"(sm.pay[1] + 2 * i));
            p = ps;
            if ((int)p >= sm.S) continue;                        /* padding (tail of group 255) */
            g = lds_u8(sm.gf + p);
            llr = __uint_as_float(lds_u32(sm.key[1] + 4 * i));
            if (dc == 4) cw = lds_u32(sm.cfg + 4 * p);
            else {
                cw = 0u;
#pragma unroll
                for (int k = 0; k < 4; k++) cw |= (k < nd ? lds_u8(sm.cfg + p * dc + d0 + k) : 1u) << (8 * k);
            }
        }
        if (g != cur) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if ((have >> k) & 1u) sts_f32(sm.key[0] + 1024 * k + 4 * ((cur ^ x[k]) & 255u), m[k] > sat[k] ? hi[k] : m[k]);
            cur = g; have = 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < nd && ((cw >> (8 * k)) & 255u) == 0u) {      /* decorrelation, :96-98 */
                m[k] = ((have >> k) & 1u) ? synd_bayes_sel(llr, m[k]) : llr;
                have |= 1u << k;
            }
        }
    }
    __syncwarp();
}
