// K4 -- fused encoder self-attention on tcgen05 / TMEM (non-causal, T <= 512 keys, head_dim 64).
//
// Replaces HF WhisperAttention.forward for the encoder (modeling_whisper.py:310-357): softmax(q k^T) v
// with q already scaled (the 1/sqrt(64) factor is folded into the q projection weights at load time).
//
// Persistent CTAs (one per SM) walk the (head, window) units; per unit a CTA loops over the 128-query tiles.  Q, K and V are
// TMA-loaded straight out of the packed [B*T, 3d] QKV activation (128-byte swizzle; keys beyond T are zero-filled by TMA and
// masked).  Only one CTA fits per SM (208 KB of shared memory, all 512 TMEM columns), so a unit's K/V load cannot hide behind
// another CTA: instead the NEXT unit's K and first Q tile are requested as soon as the last S = Q K^T of the current unit has
// completed (the K buffer is dead from then on) and its V as soon as the last P V has completed.
//   roles          : 16 softmax warps (512 threads) + one issuing warp whose lane 0 launches every TMA load and every MMA
//   S = Q K^T      : 2 x (M=128, N=256, K=64) tcgen05.mma into all 512 TMEM columns (fp32)
//   softmax        : thread r owns TMEM lane r = query row r: pass 1 row max, pass 2 exp2 + row sum,
//                    P written as bf16 into a double-buffered, manually 128B-swizzled smem tile
//   O = P V        : per 64-key chunk one (M=128, N=64, K=64) MMA group; V is consumed MN-major
//                    (head_dim contiguous) exactly as TMA delivered it.  O aliases the first 64
//                    columns of S, which are dead once chunk 0 of P has been produced.
//   epilogue       : O / rowsum -> bf16 -> out[B*T, d]
#include "common.cuh"
#include "wsb_internal.h"

namespace wsb {

#ifndef WSB_ATT_POLY_SEL
#define WSB_ATT_POLY_SEL 0x8888      // bit j set: column pair j of a thread's 16 pairs per chunk takes the polynomial exp2 (0 = all on MUFU)
#endif
constexpr int kAttThreads = 512;    // softmax threads: four warps per TMEM lane quarter, each owns a quarter of the key columns
constexpr int kAttAll = kAttThreads + 32;   // + one warp whose lane 0 issues every TMA load and every MMA
constexpr int kAttQ = 128;          // query rows per tile
constexpr int kAttKeys = 512;       // padded key count
constexpr int kHd = 64;
constexpr int kAttChunk = 128;      // keys per P tile / PV MMA group
constexpr int kAttSmem = (kAttQ * kHd + 2 * kAttKeys * kHd + 2 * kAttQ * kAttChunk) * 2 + 1024 + 128;

// barrier among the 512 softmax threads (the issuing warp never joins it)
__device__ __forceinline__ void softmax_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kAttThreads) : "memory"); }

__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    return (static_cast<unsigned long long>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
}

// K and V of a unit are loaded once and shared by its query tiles; the Q tile is reloaded as soon as its S has completed;
// TMEM (512 columns) and the mbarriers are set up once per CTA, so every barrier parity below counts over the CTA's
// lifetime: `it` = units done by this CTA, `tq` = query tiles done by this CTA.
__global__ void __launch_bounds__(kAttAll, 1)
encoder_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                         __nv_bfloat16* __restrict__ out, int T, int d, int n_heads, int n_units) {
    extern __shared__ unsigned char att_smem_raw[];
    unsigned char* smem = att_smem_raw + ((1024u - (smem_u32(att_smem_raw) & 1023u)) & 1023u);
    unsigned char* sQ = smem;                                   // 16 KB (reloaded as soon as S = QK^T has completed)
    unsigned char* sK = sQ + kAttQ * kHd * 2;                   // 64 KB
    unsigned char* sV = sK + kAttKeys * kHd * 2;                // 64 KB
    unsigned char* sP = sV + kAttKeys * kHd * 2;                // 2 x 32 KB (each: two 128x64 swizzled sub-tiles)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAttQ * kAttChunk * 2);
    uint64_t* bar_kv = bars;
    uint64_t* bar_q = bars + 1;
    uint64_t* bar_v = bars + 2;                                 // V arrives after K: S = Q K^T starts as soon as K and Q are in
    uint64_t* bar_s = bars + 3;
    uint64_t* bar_p = bars + 4;                                 // [2] P buffer consumed by the tensor core
    uint64_t* bar_o = bars + 6;
    uint64_t* bar_ready = bars + 7;                             // [2] every warp has written its part of the P buffer
    uint64_t* bar_epi = bars + 9;                               // every warp has read its part of O out of TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int quarter = warp & 3, colgrp = warp >> 2;   // TMEM lanes 32*quarter.., key-column group (0..3)
    __shared__ float s_xchg[4][kAttQ];                        // row maxima of the four key-column groups
    __shared__ float s_xsum[4][kAttQ];                        // row sums (own array: no barrier between its write and s_xchg's reads)
    const int n_qt = (T + kAttQ - 1) / kAttQ;
    int unit = blockIdx.x;                                      // grid <= n_units
    int h = unit % n_heads, b = unit / n_heads;

    if (tid == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        mbar_init(bar_kv, 1);
        mbar_init(&bar_q[0], 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_s, 1);
        mbar_init(&bar_p[0], 1);
        mbar_init(&bar_p[1], 1);
        mbar_init(bar_o, 1);
        mbar_init(&bar_ready[0], kAttThreads / 32);
        mbar_init(&bar_ready[1], kAttThreads / 32);
        mbar_init(bar_epi, kAttThreads / 32);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

#ifdef WSB_ATT_TRACE
    unsigned long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    unsigned long long t_prev = now();
#define ATT_STAMP(i) do { const unsigned long long t_ = now(); tr[i] += t_ - t_prev; t_prev = t_; } while (0)
#else
#define ATT_STAMP(i) do { } while (0)
#endif
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;
    const int row = tid & (kAttQ - 1);                     // query row within the tile == TMEM lane

    if (warp == kAttThreads / 32) {
        // ---- the issuing warp: every TMA load and every MMA of the CTA comes from its lane 0, so that no softmax thread carries
        // the ~100 serial instructions of an MMA group on the critical path of its tile ---------------------------------------
        if ((tid & 31) == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);         // B (= V) is MN-major
            mbar_arrive_expect_tx(bar_kv, kAttKeys * kHd * 2);
            tma_load_3d(sK, &tm_kv, bar_kv, d + h * kHd, 0, b);
            tma_load_3d(sK + 256 * kHd * 2, &tm_kv, bar_kv, d + h * kHd, 256, b);
            mbar_arrive_expect_tx(bar_q, kAttQ * kHd * 2);
            tma_load_3d(sQ, &tm_q, bar_q, h * kHd, 0, b);
            mbar_arrive_expect_tx(bar_v, kAttKeys * kHd * 2);
            tma_load_3d(sV, &tm_kv, bar_v, 2 * d + h * kHd, 0, b);
            tma_load_3d(sV + 256 * kHd * 2, &tm_kv, bar_v, 2 * d + h * kHd, 256, b);
            const uint64_t dq = umma_desc_k_sw128(smem_u32(sQ));
            const uint64_t dk0 = umma_desc_k_sw128(smem_u32(sK)), dk1 = umma_desc_k_sw128(smem_u32(sK + 256 * kHd * 2));
#pragma unroll 1
            for (int it = 0; unit < n_units; ++it, unit += gridDim.x) {
                h = unit % n_heads;
                b = unit / n_heads;
                const int next_unit = unit + gridDim.x;
#pragma unroll 1
                for (int qt = 0; qt < n_qt; ++qt) {
                    const int tq = it * n_qt + qt;
                    // S = Q K^T.  Upper key half (TMEM columns 256..511): for qt > 0 it was issued behind the previous
                    // tile's last P V (those columns were dead by then); lower half: once the previous O has been read out.
                    if (qt == 0) {
                        mbar_wait(bar_kv, it & 1);
                        mbar_wait(bar_q, tq & 1);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < kHd / 16; ++k) umma_bf16_ss(tmem + 256, dq + 2 * k, dk1 + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                    }
                    if (tq > 0) mbar_wait(bar_epi, (tq - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < kHd / 16; ++k) umma_bf16_ss(tmem, dq + 2 * k, dk0 + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                    umma_commit(bar_s);
                    mbar_wait(bar_s, tq & 1);                   // S is in TMEM: the Q buffer is free (and after the last tile K too)
                    if (qt + 1 < n_qt) {
                        mbar_arrive_expect_tx(bar_q, kAttQ * kHd * 2);
                        tma_load_3d(sQ, &tm_q, bar_q, h * kHd, (qt + 1) * kAttQ, b);
                    } else if (next_unit < n_units) {
                        const int nh = next_unit % n_heads, nb = next_unit / n_heads;
                        mbar_arrive_expect_tx(bar_kv, kAttKeys * kHd * 2);
                        tma_load_3d(sK, &tm_kv, bar_kv, d + nh * kHd, 0, nb);
                        tma_load_3d(sK + 256 * kHd * 2, &tm_kv, bar_kv, d + nh * kHd, 256, nb);
                        mbar_arrive_expect_tx(bar_q, kAttQ * kHd * 2);
                        tma_load_3d(sQ, &tm_q, bar_q, nh * kHd, 0, nb);
                    }
#pragma unroll 1
                    for (int c = 0; c < kAttKeys / kAttChunk; ++c) {
                        const int buf = c & 1;
                        const int use = tq * (kAttKeys / kAttChunk / 2) + (c >> 1);
                        mbar_wait(&bar_ready[buf], use & 1);
                        if (qt == 0 && c == 0) mbar_wait(bar_v, it & 1);
                        tc_fence_after();
                        const uint64_t dv = umma_desc_mn_sw128(smem_u32(sV + c * kAttChunk * kHd * 2));
#pragma unroll
                        for (int k = 0; k < kAttChunk / 16; ++k) {     // 16 keys per MMA: P advances 32 B (next sub-tile after
                                                                       // 64 keys), V advances 16 rows = 2048 B
                            const uint64_t dp = umma_desc_k_sw128(smem_u32(sP + buf * kAttQ * kAttChunk * 2 + (k >> 2) * kAttQ * kHd * 2));
                            umma_bf16_ss(tmem, dp + 2 * (k & 3), dv + ((16 * kHd * 2) >> 4) * k, idesc_o, (c > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(&bar_p[buf]);
                    }
                    umma_commit(bar_o);
                    if (qt + 1 < n_qt) {
                        // every warp has read its columns of the last chunk (bar_ready): columns 256..511 are dead
                        mbar_wait(bar_q, (tq + 1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < kHd / 16; ++k) umma_bf16_ss(tmem + 256, dq + 2 * k, dk1 + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                    } else if (next_unit < n_units) {           // V is dead once the last P V has completed: the next unit's V
                        mbar_wait(bar_o, tq & 1);
                        const int nh = next_unit % n_heads, nb = next_unit / n_heads;
                        mbar_arrive_expect_tx(bar_v, kAttKeys * kHd * 2);
                        tma_load_3d(sV, &tm_kv, bar_v, 2 * d + nh * kHd, 0, nb);
                        tma_load_3d(sV + 256 * kHd * 2, &tm_kv, bar_v, 2 * d + nh * kHd, 256, nb);
                    }
                }
            }
        }
        __syncwarp();                                           // lanes 1..31 wait here for lane 0: the warp reaches the final barrier converged
    } else {
#pragma unroll 1
    for (int it = 0; unit < n_units; ++it, unit += gridDim.x) {
    h = unit % n_heads;
    b = unit / n_heads;
#pragma unroll 1
    for (int qt = 0; qt < n_qt; ++qt) {
        const int tq = it * n_qt + qt;                          // query tiles this CTA has finished before this one
        ATT_STAMP(0);
        mbar_wait(bar_s, tq & 1);
        tc_fence_after();
        ATT_STAMP(1);                                       // S = Q K^T complete

        // pass 1: row max over the T valid keys (each warp of the pair scans its 256-column half)
        float rmax = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < kAttKeys / 128; ++c) {
            uint32_t r[32];
            const int n0 = colgrp * 128 + c * 32;
            tmem_ld_32x32(lane_base + n0, r);
            tmem_ld_wait();
            if (n0 + 32 <= T) {
#pragma unroll
                for (int i = 0; i < 32; ++i) rmax = fmaxf(rmax, __uint_as_float(r[i]));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (n0 + i < T) rmax = fmaxf(rmax, __uint_as_float(r[i]));
            }
        }
        s_xchg[colgrp][row] = rmax;
        softmax_sync();
        rmax = fmaxf(fmaxf(s_xchg[0][row], s_xchg[1][row]), fmaxf(s_xchg[2][row], s_xchg[3][row]));
        const float mscaled = rmax * kLog2e;
        ATT_STAMP(2);                                       // pass 1 (row max)

        // pass 2: P chunks + PV MMAs.  The TMEM columns of chunk c + 1 are requested before chunk c is exponentiated, the
        // scale and the row sum run as packed f32x2 instructions (FFMA2 / FADD2), and the warps hand their part of the P
        // buffer over with an mbarrier arrive instead of a CTA-wide barrier (only the issuing warp waits for the slowest).
        unsigned long long rsum2 = 0ull;                       // (sum of even columns, sum of odd columns)
        uint32_t rnext[32];
        tmem_ld_32x32(lane_base + colgrp * 32, rnext);
        const float nm = -mscaled;
        const unsigned long long scale2 = pack_f32x2(kLog2e, kLog2e), shift2 = pack_f32x2(nm, nm);
        // one column pair in four (WSB_ATT_POLY_SEL) takes its exponential as a degree-3 polynomial on the FMA / integer pipes
        // (relative error 8e-4, below the bf16 rounding of P): a polynomial pair costs ~14 issue slots, a MUFU pair 2 issue
        // slots but 16 cycles of the 16-lane MUFU pipe, and the pass is issue-bound -- measured 0.747 / 0.735 / 0.729 / 0.728 ms
        // per launch with 8 / 6 / 5 / 4 of the 16 pairs on the polynomial
        constexpr unsigned kPolySel = WSB_ATT_POLY_SEL;
        const unsigned long long kMagic2 = pack_f32x2(12582912.0f, 12582912.0f), kNegMagic2 = pack_f32x2(-12582912.0f, -12582912.0f);
        const unsigned long long kNegOne2 = pack_f32x2(-1.0f, -1.0f), kOne2 = pack_f32x2(1.0f, 1.0f);
        const unsigned long long kC3 = pack_f32x2(0.05550411f, 0.05550411f), kC2 = pack_f32x2(0.24022651f, 0.24022651f);
        const unsigned long long kC1 = pack_f32x2(0.69314718f, 0.69314718f);
#pragma unroll
        for (int c = 0; c < kAttKeys / kAttChunk; ++c) {
            const int buf = c & 1;
            const int use = tq * (kAttKeys / kAttChunk / 2) + (c >> 1);  // how many times this buffer was used before
            // this warp's 32 of the chunk's 128 keys: sub-tile (64 keys) colgrp>>1, 32-key half colgrp&1
            unsigned char* pbuf = sP + buf * kAttQ * kAttChunk * 2 + (colgrp >> 1) * kAttQ * kHd * 2 + row * 128;
            {
                const int hlf = colgrp & 1;
                uint32_t r[32];
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = rnext[i];
                if (c + 1 < kAttKeys / kAttChunk) tmem_ld_32x32(lane_base + (c + 1) * kAttChunk + colgrp * 32, rnext);
                const int n0 = c * kAttChunk + colgrp * 32;
                float pv[32];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    unsigned long long x2;
                    const unsigned long long s2 = (static_cast<unsigned long long>(r[i + 1]) << 32) | r[i];
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x2) : "l"(s2), "l"(scale2), "l"(shift2));
                    float e0, e1;
                    if ((kPolySel >> (i >> 1)) & 1u) {
                        // this pair on the FMA / integer pipes: 2^x = 2^n * p(f), n = round(x), f = x - n in [-0.5, 0.5]
                        const float a0 = fmaxf(__uint_as_float(static_cast<uint32_t>(x2)), -125.0f);
                        const float a1 = fmaxf(__uint_as_float(static_cast<uint32_t>(x2 >> 32)), -125.0f);
                        const unsigned long long a2 = pack_f32x2(a0, a1);
                        unsigned long long t2, n2, f2, p2;
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t2) : "l"(a2), "l"(kMagic2));
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(n2) : "l"(t2), "l"(kNegMagic2));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f2) : "l"(n2), "l"(kNegOne2), "l"(a2));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(kC3), "l"(f2), "l"(kC2));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(p2), "l"(f2), "l"(kC1));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(p2), "l"(f2), "l"(kOne2));
                        e0 = __int_as_float(static_cast<int>(static_cast<uint32_t>(p2)) + ((static_cast<int>(static_cast<uint32_t>(t2)) - 0x4B400000) << 23));
                        e1 = __int_as_float(static_cast<int>(static_cast<uint32_t>(p2 >> 32)) + ((static_cast<int>(static_cast<uint32_t>(t2 >> 32)) - 0x4B400000) << 23));
                    } else {
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(__uint_as_float(static_cast<uint32_t>(x2))));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(__uint_as_float(static_cast<uint32_t>(x2 >> 32))));
                    }
                    pv[i] = e0;
                    pv[i + 1] = e1;
                }
                if (n0 + 32 > T) {                             // warp-uniform: only the warps that straddle T mask anything
                    asm volatile("" ::: "memory");             // (keeps this a branch: if-converted it costs every warp 64 instructions per chunk)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n0 + i >= T) pv[i] = 0.0f;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const unsigned long long p2 = pack_f32x2(pv[i], pv[i + 1]);
                    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(rsum2) : "l"(p2));
                }
                if (use > 0) mbar_wait(&bar_p[buf], (use - 1) & 1);     // the tensor core is done with this P buffer
#pragma unroll
                for (int j = 0; j < 4; ++j) {                  // four 16-byte chunks of this 32-key half
                    uint4 pk;
                    pk.x = pack_bf16x2(pv[8 * j + 0], pv[8 * j + 1]);
                    pk.y = pack_bf16x2(pv[8 * j + 2], pv[8 * j + 3]);
                    pk.z = pack_bf16x2(pv[8 * j + 4], pv[8 * j + 5]);
                    pk.w = pack_bf16x2(pv[8 * j + 6], pv[8 * j + 7]);
                    const int chunk = hlf * 4 + j;             // 16-byte chunk index within the 128-byte row
                    *reinterpret_cast<uint4*>(pbuf + ((chunk ^ (row & 7)) << 4)) = pk;
                }
            }
            fence_proxy_async_smem();                          // generic-proxy smem writes -> visible to the MMA
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&bar_ready[buf]);          // the issuing warp launches this chunk's P V when all 16 have arrived
        }
        const float rsum = __uint_as_float(static_cast<uint32_t>(rsum2)) + __uint_as_float(static_cast<uint32_t>(rsum2 >> 32));
        ATT_STAMP(3);                                       // pass 2 (exp, P, PV issue)
        mbar_wait(bar_o, tq & 1);
        tc_fence_after();
        ATT_STAMP(4);                                       // last PV MMAs complete

        // epilogue: O / rowsum (the row sums of the four key-column groups are combined through shared memory); every warp
        // takes 16 of the 64 output dims of its 32 rows
        s_xsum[colgrp][row] = rsum;
        softmax_sync();
        const int q_row = qt * kAttQ + row;
        const float inv = 1.0f / ((s_xsum[0][row] + s_xsum[1][row]) + (s_xsum[2][row] + s_xsum[3][row]));
        {
            uint32_t r[16];
            tmem_ld_32x16(lane_base + colgrp * 16, r);
            tmem_ld_wait();
            if (q_row < T) {
                __nv_bfloat16* o = out + (static_cast<size_t>(b) * T + q_row) * d + h * kHd + colgrp * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 8) {
                    uint4 pk;
                    pk.x = pack_bf16x2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
                    pk.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
                    pk.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
                    pk.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
                    *reinterpret_cast<uint4*>(o + i) = pk;
                }
            }
        }
        // the lower half of the next tile's S overwrites the TMEM columns O was just read from: tell the issuing warp
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bar_epi);
        ATT_STAMP(5);                                       // epilogue
    }
    }
    }
    tc_fence_before();
    __syncthreads();                                            // every MMA has completed (bar_o) and every TMEM read is done
#ifdef WSB_ATT_TRACE
    if (tid == 0 && blockIdx.x == 3)
        printf("attention trace (ns, thread 0, all units of CTA 3): wait-loads+issue S %llu | S mma %llu | pass1 %llu | pass2 %llu | PV tail %llu | epilogue %llu\n",
               tr[0], tr[1], tr[2], tr[3], tr[4], tr[5]);
#endif
    if (warp == 1) tmem_dealloc<512>(tmem);
}

int encoder_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int n_heads, cudaStream_t stream) {
    WSB_REQUIRE(T <= kAttKeys && T > 0, "encoder attention supports up to 512 positions");
    if (B <= 0) return 0;
    const int d = n_heads * kHd;
    static PerDeviceOnce once;
    int dev = 0;
    if (once.need(&dev)) {
        WSB_CHECK_CUDA(cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem));
        once.mark(dev);
    }
    CUtensorMap tm_q, tm_kv;
    uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(T) * 3 * d * 2};
    uint32_t box_q[3] = {static_cast<uint32_t>(kHd), static_cast<uint32_t>(kAttQ), 1};
    uint32_t box_kv[3] = {static_cast<uint32_t>(kHd), 256, 1};
    int rc = make_tmap_bf16(&tm_q, qkv, 3, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_bf16(&tm_kv, qkv, 3, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const int n_units = n_heads * B;
    int num_sms = 148;
    WSB_CHECK_CUDA(cudaGetDevice(&dev));
    WSB_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = std::min(n_units, std::max(8, num_sms - g_sm_reserve));
    encoder_attention_kernel<<<grid, kAttAll, kAttSmem, stream>>>(tm_q, tm_kv, out, T, d, n_heads, n_units);
    WSB_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace wsb
