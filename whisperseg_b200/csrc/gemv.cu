// K5c -- skinny linear layers for decode batches of at most 64 rows (the tail of a greedy decode after batch
// compaction, or a short recording with a handful of windows).
//
// At few rows the tcgen05 split-K GEMM + second-phase reduce pair is pure fixed latency (TMEM allocation,
// tensor-map fetch, pipeline fill, plane write + re-read, two launches): ~13 us per linear layer against
// ~2 us of weight streaming.  This kernel does the whole layer in one launch of <= 148 CTAs (one per SM):
//   * a CTA owns 8*NT output features and the full K extent -> no split-K, no partial planes.  Its weight
//     tile is fetched by TMA bulk copies (one per weight row, padded destination rows so that the fragment
//     reads are bank-conflict free), all issued at kernel entry on one mbarrier: maximum memory-level
//     parallelism at zero register cost;
//   * the consumer's LayerNorm is fused in WITHOUT a separate pass: the producer of the residual stream (the
//     previous residual-update launch of this kernel, or the row-statistics kernel after the embedding) leaves
//     per-CTA partial (sum, sum of squares) of every row; the consumer adds them in a fixed order and either
//     (IN_LN 1) normalises exactly the fp32 elements each lane needs for its MMA fragments, or (IN_LN 2, the
//     default) multiplies the producer's bf16 copy of the rows with weights that have the LayerNorm affine
//     folded in and applies rstd (acc - mean c1) + c2 in the epilogue (weights.py: fold_layernorm);
//   * 16 rows are exactly the M of a warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  MT = 1, 2 or 4
//     m-tiles: the 16 warps form MT groups, a group owns 16 rows and its warps split K (the k order inside a
//     32-element block is permuted identically for both operands so that every lane's 8 contiguous elements
//     feed two MMAs); the partial tiles are summed through shared memory and the epilogue (bias; GELU -> bf16 |
//     fp32 | in-place fp32 residual add + bf16 copy + row statistics for the next LayerNorm) is applied.
// HBM-bound by design at <= 16 rows (it streams each weight byte once; the decode step is launch-latency-bound
// around it); beyond that every CTA re-reads all activations and the kernel becomes L2-bound.
// (tcgen05 needs M = 128 tiles and a TMEM prologue -- the wrong tool at M <= 16; this is deliberate.)
#include "common.cuh"
#include "wsb_internal.h"

#include <algorithm>

namespace wsb {

// -DWSB_GV_THREADS=256 builds the two-CTAs-per-SM variant (half the registers and shared memory per CTA, so that under
// programmatic dependent launch launch i+1 is resident next to launch i).  Measured in round 2 and rejected: the bench
// decode went from 414 to 464 ms (profiles/r2_decode_gemv_256_vs_512_threads.txt) -- 8 warps keep half as many
// activation loads in flight and the smaller tiles need more CTAs; 512 threads, one CTA per SM stays the default.
#ifndef WSB_GV_THREADS
#define WSB_GV_THREADS 512
#endif
constexpr int kGvThreads = WSB_GV_THREADS;
constexpr int kGvMinBlocks = kGvThreads <= 256 ? 2 : 1;
constexpr int kGvWarps = kGvThreads / 32;
constexpr int kGvBatch = 3;             // k-blocks (of 32) per warp whose activation fragments are in flight
constexpr int kGvMaxNT = 5;
constexpr int kGvWPad = 64;             // bytes of padding per weight row in smem (row shift = 16 banks)
constexpr int kGvMaxRows = 64;          // 4 m-tiles of 16 rows
constexpr int kGvMaxSmem = (kGvMinBlocks == 2 ? 112 : 220) * 1024;

struct GvParams {
    const float* x;                     // LN input: fp32 [M][K] (IN_LN) ...
    const float* stats;                 // [parts][MP][2] partial (sum, sum of squares) of every row of x (MP = 16, 32 or 64)
    int stats_parts;
    const float* gamma;
    const float* beta;
    const __nv_bfloat16* a;             // ... or bf16 activations [M][K]
    const float* c1;                    // IN_LN 2 (folded LayerNorm): a = bf16(x), W = bf16(W o gamma), c1[n] = sum_k W[n][k],
                                        // bias = b + W beta; out = rstd (acc - mean c1) + bias, statistics as for IN_LN 1
    const __nv_bfloat16* W;             // [N][K]
    const float* bias;                  // [N] or null
    float* out_f32;                     // EPI 0: [M][N]
    __nv_bfloat16* out_bf16;            // EPI 1: [M][N] = gelu(.)
    float* resid;                       // EPI 2: [M][N] += .
    float* stats_out;                   // EPI 2: [gridDim.x][MP][2] partial row statistics of the updated rows
    __nv_bfloat16* xb_out;              // EPI 2, optional: bf16 copy of the updated rows (input of a folded-LayerNorm consumer)
    const unsigned char* row_skip;      // [M] or null: rows not stored
    int* fold_flag;                     // IN_LN 2, optional: set to 1 when a live row has |mean| > 2 std (folded-LayerNorm guard)
    int M, N, K;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint4 ln_pack8(const float4& v0, const float4& v1, float mean, float rstd, const float* gam,
                                          const float* bet) {
    const float4 g0 = *reinterpret_cast<const float4*>(gam), g1 = *reinterpret_cast<const float4*>(gam + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(bet), b1 = *reinterpret_cast<const float4*>(bet + 4);
    uint4 r;
    r.x = pack_bf16x2((v0.x - mean) * rstd * g0.x + b0.x, (v0.y - mean) * rstd * g0.y + b0.y);
    r.y = pack_bf16x2((v0.z - mean) * rstd * g0.z + b0.z, (v0.w - mean) * rstd * g0.w + b0.w);
    r.z = pack_bf16x2((v1.x - mean) * rstd * g1.x + b1.x, (v1.y - mean) * rstd * g1.y + b1.y);
    r.w = pack_bf16x2((v1.z - mean) * rstd * g1.z + b1.z, (v1.w - mean) * rstd * g1.w + b1.w);
    return r;
}

// MT = number of 16-row m-tiles (1, 2 or 4 -> up to 64 rows).  The 16 warps are split into MT groups of
// KS = 16 / MT warps: group mt owns rows [16 mt, 16 mt + 16) and its warps split K.  The weight tile is
// fetched once per CTA whatever MT is; the activation fragments of a warp are loaded in batches of kGvBatch
// k-blocks, the next batch in flight while the current one feeds the MMAs.  m-tiles whose rows are all
// finished are skipped (no loads, no MMAs), finished rows of a live tile are not loaded.
template <int IN_LN, int EPI, int NT, int MT>
__global__ void __launch_bounds__(kGvThreads, kGvMinBlocks) gemv16_kernel(const GvParams p) {
    constexpr int KS = kGvWarps / MT;                    // warps splitting K inside one m-tile
    constexpr int MP = 16 * MT;                          // padded row count (statistics layout [parts][MP][2])
    extern __shared__ __align__(128) unsigned char gv_smem[];
    __shared__ float s_mean[MP], s_rstd[MP];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int mt = warp / KS, kw = warp % KS, rb = mt * 16;
    const int n0 = blockIdx.x * 8 * NT;
    const int nkb = p.K >> 5;
    const int w_stride = p.K * 2 + kGvWPad;              // bytes
    unsigned char* w_s = gv_smem;
    float* gam_s = reinterpret_cast<float*>(gv_smem + static_cast<size_t>(8 * NT) * w_stride);
    float* bet_s = gam_s + (IN_LN == 1 ? p.K : 0);
    float* red_s = bet_s + (IN_LN == 1 ? p.K : 0);            // [16 warps][NT][16][8]
    float* new_s = red_s + kGvWarps * NT * 128;          // EPI 2: [NT][MP][8]
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    // Everything before pdl_wait() touches constants only (weights, LayerNorm affine): under programmatic
    // dependent launch it overlaps the tail of the previous kernel.
    if (warp == 0) {                                     // weight tile: one bulk copy per row, all in flight at once
        const int rows = min(8 * NT, p.N - n0);
        if (lane == 0) mbar_arrive_expect_tx(&bar, static_cast<uint32_t>(rows) * p.K * 2);
        __syncwarp();
        for (int r = lane; r < rows; r += 32)
            bulk_copy_g2s(w_s + static_cast<size_t>(r) * w_stride, p.W + static_cast<long long>(n0 + r) * p.K,
                          static_cast<uint32_t>(p.K) * 2, &bar);
    }

    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    if constexpr (IN_LN == 1) {
        for (int j = tid; j < (p.K >> 2); j += kGvThreads) {
            reinterpret_cast<float4*>(gam_s)[j] = __ldg(reinterpret_cast<const float4*>(p.gamma) + j);
            reinterpret_cast<float4*>(bet_s)[j] = __ldg(reinterpret_cast<const float4*>(p.beta) + j);
        }
    }
    // epilogue operands of this thread's output elements: the constants (bias, c1) are fetched before the dependency
    // wait, the residual rows right after it -- every global load of the launch is in flight at once and the kernel
    // pays ONE L2 round trip (~1 us), not one per stage (measured with mega.cu's stage trace)
    constexpr int kEpi = (MT * NT * 128 + kGvThreads - 1) / kGvThreads;
    float pre_bias[kEpi], pre_c1[IN_LN == 2 ? kEpi : 1], pre_res[EPI == 2 ? kEpi : 1];
#pragma unroll
    for (int e = 0; e < kEpi; ++e) {
        const int idx = tid + e * kGvThreads;
        const int rem = idx % (NT * 128);
        const int n = n0 + (rem >> 7) * 8 + (rem & 7);
        const bool ok = idx < MT * NT * 128 && n < p.N;
        pre_bias[e] = (ok && p.bias) ? __ldg(p.bias + n) : 0.0f;
        if constexpr (IN_LN == 2) pre_c1[e] = ok ? __ldg(p.c1 + n) : 0.0f;
    }
    pdl_wait();

    // A fragments are loaded for every row inside M -- also finished ones, whose results are never stored: gating the
    // loads on the finished flags would put a dependent round trip in front of them.  An m-tile whose rows have all
    // finished is still skipped (no MMAs, no further loads).
    const int r_lo = rb + g, r_hi = rb + g + 8;
    const bool use_lo = r_lo < p.M, use_hi = r_hi < p.M;
    const int all_blocks = kw < nkb ? (nkb - kw + KS - 1) / KS : 0;                   // k-blocks kw, kw + KS, ...
    int my_blocks = all_blocks;
    int n_batches = (my_blocks + kGvBatch - 1) / kGvBatch;
    unsigned char skip_lo = 0, skip_hi = 0;
    if (p.row_skip) {
        if (use_lo) skip_lo = p.row_skip[r_lo];
        if (use_hi) skip_hi = p.row_skip[r_hi];
    }
    if constexpr (EPI == 2) {
#pragma unroll
        for (int e = 0; e < kEpi; ++e) {
            const int idx = tid + e * kGvThreads;
            const int m = idx / (NT * 128), rem = idx - m * (NT * 128);
            const int row = m * 16 + ((rem & 127) >> 3);
            const int n = n0 + (rem >> 7) * 8 + (rem & 7);
            pre_res[e] = (idx < MT * NT * 128 && row < p.M && n < p.N) ? p.resid[static_cast<long long>(row) * p.N + n] : 0.0f;
        }
    }

    // raw activation fragments of one batch: fp32 (LayerNorm path) or bf16
    float4 xl[IN_LN == 1 ? kGvBatch : 1][2], xh[IN_LN == 1 ? kGvBatch : 1][2];
    uint4 a_lo[kGvBatch], a_hi[kGvBatch], n_lo[IN_LN == 1 ? 1 : kGvBatch], n_hi[IN_LN == 1 ? 1 : kGvBatch];
    auto load_batch = [&](int bi) {
#pragma unroll
        for (int i = 0; i < kGvBatch; ++i) {
            const int j = bi * kGvBatch + i;
            const int ko = (kw + j * KS) * 32 + t * 8;
            if constexpr (IN_LN == 1) {
                xl[i][0] = xl[i][1] = xh[i][0] = xh[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < my_blocks) {
                    if (use_lo) {
                        const float4* s = reinterpret_cast<const float4*>(p.x + static_cast<long long>(r_lo) * p.K + ko);
                        xl[i][0] = s[0];
                        xl[i][1] = s[1];
                    }
                    if (use_hi) {
                        const float4* s = reinterpret_cast<const float4*>(p.x + static_cast<long long>(r_hi) * p.K + ko);
                        xh[i][0] = s[0];
                        xh[i][1] = s[1];
                    }
                }
            } else {
                n_lo[i] = zero4;
                n_hi[i] = zero4;
                if (j < my_blocks) {
                    if (use_lo) n_lo[i] = *reinterpret_cast<const uint4*>(p.a + static_cast<long long>(r_lo) * p.K + ko);
                    if (use_hi) n_hi[i] = *reinterpret_cast<const uint4*>(p.a + static_cast<long long>(r_hi) * p.K + ko);
                }
            }
        }
    };
    if (n_batches > 0) load_batch(0);
    {
        const bool tile_live = __any_sync(0xffffffffu, (use_lo && !skip_lo) || (use_hi && !skip_hi));
        if (!tile_live) {
            my_blocks = 0;
            n_batches = 0;
        }
    }

    float m_lo = 0.f, rs_lo = 0.f, m_hi = 0.f, rs_hi = 0.f;
    if constexpr (IN_LN != 0) {
        // row statistics: 512 / MP threads per row add the producer's partials in a fixed order (all their loads
        // in flight together), then an xor tree inside the row's lane group
        {
            constexpr int TPR = kGvThreads / MP;         // 32, 16 or 8 threads per row
            const int r = tid / TPR, sub = tid % TPR;
            float sm = 0.0f, sq = 0.0f;
            if (r < p.M) {
                // 8 partials per thread in flight at a time, added in ascending order (adding the zero of a missing
                // partial changes nothing, so the sum is the sequential one)
                for (int base = sub; base < p.stats_parts; base += 8 * TPR) {
                    float2 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int pp = base + i * TPR;
                        v[i] = pp < p.stats_parts ? *reinterpret_cast<const float2*>(p.stats + (static_cast<long long>(pp) * MP + r) * 2)
                                                  : make_float2(0.0f, 0.0f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        sm += v[i].x;
                        sq += v[i].y;
                    }
                }
            }
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) {
                sm += __shfl_xor_sync(0xffffffffu, sm, o);
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            if (sub == 0) {
                const float mean = sm / p.K;
                const float var = fmaxf(sq / p.K - mean * mean, 0.0f);
                s_mean[r] = mean;
                s_rstd[r] = rsqrtf(var + 1e-5f);
                if constexpr (IN_LN == 2) {
                    // Folded LayerNorm rounds x, not LN(x), to bf16: its error relative to the exact path grows like
                    // sqrt(1 + mean^2 / var).  A row whose common mode dominates its spread raises the flag and the
                    // engine falls back to the exact on-the-fly LayerNorm (engine.cu: fold guard).
                    if (p.fold_flag != nullptr && blockIdx.x == 0 && r < p.M && !(p.row_skip && p.row_skip[r]) &&
                        mean * mean > 4.0f * var)
                        *p.fold_flag = 1;
                }
            }
        }
        __syncthreads();                                 // also publishes gam_s / bet_s
        m_lo = s_mean[r_lo];
        rs_lo = s_rstd[r_lo];
        m_hi = s_mean[r_hi];
        rs_hi = s_rstd[r_hi];
    }

    bool w_ready = false;
    for (int bi = 0; bi < n_batches; ++bi) {
        // raw -> bf16 fragments of this batch, then put the next batch's loads in flight
#pragma unroll
        for (int i = 0; i < kGvBatch; ++i) {
            if constexpr (IN_LN == 1) {
                const int j = bi * kGvBatch + i;
                const int ko = (kw + j * KS) * 32 + t * 8;
                a_lo[i] = zero4;
                a_hi[i] = zero4;
                if (j < my_blocks) {
                    if (use_lo) a_lo[i] = ln_pack8(xl[i][0], xl[i][1], m_lo, rs_lo, gam_s + ko, bet_s + ko);
                    if (use_hi) a_hi[i] = ln_pack8(xh[i][0], xh[i][1], m_hi, rs_hi, gam_s + ko, bet_s + ko);
                }
            } else {
                a_lo[i] = n_lo[i];
                a_hi[i] = n_hi[i];
            }
        }
        if (bi + 1 < n_batches) load_batch(bi + 1);
        if (!w_ready) {
            mbar_wait(&bar, 0);
            w_ready = true;
        }
#pragma unroll
        for (int i = 0; i < kGvBatch; ++i) {
            const int j = bi * kGvBatch + i;
            if (j < my_blocks) {
                const int kb = kw + j * KS;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint4 b = *reinterpret_cast<const uint4*>(w_s + static_cast<size_t>(nt * 8 + g) * w_stride + (kb * 32 + t * 8) * 2);
                    // lane elements [0..3] -> k slots (2t, 2t+1 | 2t+8, 2t+9) of MMA 1, elements [4..7] of MMA 2
                    mma_bf16_16816(acc[nt], a_lo[i].x, a_hi[i].x, a_lo[i].y, a_hi[i].y, b.x, b.y);
                    mma_bf16_16816(acc[nt], a_lo[i].z, a_hi[i].z, a_lo[i].w, a_hi[i].w, b.z, b.w);
                }
            }
        }
    }
    if (!w_ready) mbar_wait(&bar, 0);                    // never leave bulk copies in flight behind an exited CTA
    float* red_w = red_s + warp * (NT * 128);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        red_w[nt * 128 + g * 8 + 2 * t] = acc[nt][0];
        red_w[nt * 128 + g * 8 + 2 * t + 1] = acc[nt][1];
        red_w[nt * 128 + (g + 8) * 8 + 2 * t] = acc[nt][2];
        red_w[nt * 128 + (g + 8) * 8 + 2 * t + 1] = acc[nt][3];
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kEpi; ++e) {
        const int idx = tid + e * kGvThreads;
        if (idx >= MT * NT * 128) break;
        const int m = idx / (NT * 128), rem = idx - m * (NT * 128);
        const int nt = rem >> 7, r = (rem & 127) >> 3, c = rem & 7;
        const int row = m * 16 + r;
        const int n = n0 + nt * 8 + c;
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < KS; ++w) v += red_s[(m * KS + w) * (NT * 128) + nt * 128 + r * 8 + c];
        const bool valid = row < p.M && n < p.N;
        if constexpr (IN_LN == 2) {
            if (valid) v = s_rstd[row] * (v - s_mean[row] * pre_c1[e]);
        }
        if (valid && p.bias) v += pre_bias[e];
        const bool store = valid && !(p.row_skip && p.row_skip[row]);
        const long long o = static_cast<long long>(row) * p.N + n;
        if constexpr (EPI == 0) {
            if (store) p.out_f32[o] = v;
        }
        if constexpr (EPI == 1) {
            if (store) p.out_bf16[o] = __float2bfloat16(gelu_fast(v));
        }
        if constexpr (EPI == 2) {
            float xn = 0.0f;
            if (valid) {
                xn = pre_res[e] + (store ? v : 0.0f);
                if (store) {
                    p.resid[o] = xn;
                    if (p.xb_out) p.xb_out[o] = __float2bfloat16(xn);
                }
            }
            new_s[(nt * MP + row) * 8 + c] = xn;
        }
    }
    if constexpr (EPI == 2) {
        if (p.stats_out) {
            __syncthreads();
            // one (sum, sum of squares) partial per 8-feature n-tile and row -- independent of how many n-tiles a CTA
            // owns, so every producer of the residual stream (this kernel, mega.cu) leaves the same partials
            for (int idx = tid; idx < NT * MP; idx += kGvThreads) {
                const int nt = idx / MP, row = idx - nt * MP;
                if (n0 + nt * 8 < p.N) {
                    float sm = 0.0f, sq = 0.0f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float v = new_s[(nt * MP + row) * 8 + c];
                        sm += v;
                        sq = fmaf(v, v, sq);
                    }
                    *reinterpret_cast<float2*>(p.stats_out + (static_cast<long long>(n0 / 8 + nt) * MP + row) * 2) = make_float2(sm, sq);
                }
            }
        }
    }
}

// exact (sum, sum of squares) of every row: the statistics of a residual stream that no gemv16 launch produced
__global__ void row_stats_kernel(const float* __restrict__ x, int K, float* __restrict__ stats, __nv_bfloat16* __restrict__ xb) {
    const int r = blockIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    float sm = 0.0f, sq = 0.0f;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const float v = x[static_cast<long long>(r) * K + j];
        if (xb) xb[static_cast<long long>(r) * K + j] = __float2bfloat16(v);
        sm += v;
        sq = fmaf(v, v, sq);
    }
    __shared__ float a[8], b[8];
    sm = warp_sum(sm);
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) {
        a[threadIdx.x >> 5] = sm;
        b[threadIdx.x >> 5] = sq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s0 = 0.0f, s1 = 0.0f;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) {
            s0 += a[w];
            s1 += b[w];
        }
        stats[r * 2] = s0;
        stats[r * 2 + 1] = s1;
    }
}

int row_stats16(const float* x, int M, int K, float* stats, cudaStream_t stream, __nv_bfloat16* xb) {
    WSB_REQUIRE(M >= 1 && M <= kGvMaxRows, "row_stats16 handles at most 64 rows");
    WSB_CHECK_CUDA(launch_kernel(row_stats_kernel, dim3(M), dim3(256), 0, stream, x, K, stats, xb));
    count_launch();
    return 0;
}

int row_stats_any(const float* x, int M, int K, float* stats, __nv_bfloat16* xb, cudaStream_t stream) {
    if (M <= 0) return 0;
    WSB_CHECK_CUDA(launch_kernel(row_stats_kernel, dim3(M), dim3(256), 0, stream, x, K, stats, xb));
    count_launch();
    return 0;
}

static size_t gv_smem_bytes(int nt, int K, int in_ln, int epi, int mt) {
    return static_cast<size_t>(8 * nt) * (static_cast<size_t>(K) * 2 + kGvWPad) + (in_ln == 1 ? static_cast<size_t>(K) * 8 : 0) +
           sizeof(float) * kGvWarps * nt * 128 + (epi == 2 ? sizeof(float) * nt * 16 * mt * 8 : 0);
}
// n-tiles (of 8 output features) per CTA: about one CTA per SM, fewer features per CTA when the tile would not fit
static int gv_pick_nt(int N, int K, int in_ln, int epi, int mt) {
    int nt = std::min(kGvMaxNT, std::max(1, ceil_div(ceil_div(N, 8), 148)));
    while (nt > 1 && gv_smem_bytes(nt, K, in_ln, epi, mt) > static_cast<size_t>(kGvMaxSmem)) --nt;
    return nt;
}
int gemv16_parts(int N, int K) {                         // partial statistics a residual-update launch leaves: one per 8 features
    (void)K;
    return ceil_div(N, 8);
}
int gemv16_max_rows() { return kGvMaxRows; }

template <int IN_LN, int EPI, int NT, int MT>
static int launch_gemv(const GvParams& p, cudaStream_t stream) {
    const size_t smem = gv_smem_bytes(NT, p.K, IN_LN, EPI, MT);
    WSB_REQUIRE(smem <= kGvMaxSmem, "gemv16: weight tile does not fit in shared memory (K too large)");
    static PerDeviceOnce once;
    int dev = 0;
    if (once.need(&dev)) {
        WSB_CHECK_CUDA(cudaFuncSetAttribute(gemv16_kernel<IN_LN, EPI, NT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGvMaxSmem));
        once.mark(dev);
    }
    WSB_CHECK_CUDA(launch_kernel(gemv16_kernel<IN_LN, EPI, NT, MT>, dim3(ceil_div(p.N, 8 * NT)), dim3(kGvThreads), smem, stream, p));
    count_launch();
    return 0;
}

template <int IN_LN, int EPI, int MT>
static int dispatch_nt(const GvParams& p, int nt, cudaStream_t stream) {
    switch (nt) {
        case 1: return launch_gemv<IN_LN, EPI, 1, MT>(p, stream);
        case 2: return launch_gemv<IN_LN, EPI, 2, MT>(p, stream);
        case 3: return launch_gemv<IN_LN, EPI, 3, MT>(p, stream);
        case 4: return launch_gemv<IN_LN, EPI, 4, MT>(p, stream);
        default: return launch_gemv<IN_LN, EPI, 5, MT>(p, stream);
    }
}

template <int IN_LN, int EPI>
static int dispatch_mt(const GvParams& p, int nt, cudaStream_t stream) {
    if (p.M <= 16) return dispatch_nt<IN_LN, EPI, 1>(p, nt, stream);
    if (p.M <= 32) return dispatch_nt<IN_LN, EPI, 2>(p, nt, stream);
    return dispatch_nt<IN_LN, EPI, 4>(p, nt, stream);
}

int gemv16(const Gemv16Args& a, cudaStream_t stream) {
    WSB_REQUIRE(a.M >= 1 && a.M <= kGvMaxRows, "gemv16 handles at most 64 rows");
    WSB_REQUIRE(a.K % 32 == 0 && a.K >= 32, "gemv16: K must be a multiple of 32");
    WSB_REQUIRE((a.x != nullptr) != (a.a != nullptr), "gemv16: exactly one of the fp32 (LayerNorm) and bf16 inputs");
    WSB_REQUIRE(!a.x || (a.K <= 1536 && a.gamma && a.beta && a.stats && a.stats_parts >= 1),
                "gemv16: fused LayerNorm needs gamma/beta, row statistics and K <= 1536");
    WSB_REQUIRE(!a.c1 || (a.a && a.stats && a.stats_parts >= 1 && !a.resid), "gemv16: folded LayerNorm needs bf16 rows and row statistics");
    const int outs = (a.out_f32 != nullptr) + (a.out_bf16_gelu != nullptr) + (a.resid != nullptr);
    WSB_REQUIRE(outs == 1, "gemv16: exactly one output mode");
    WSB_REQUIRE(ceil_div(a.N, 8 * kGvMaxNT) <= 1024, "gemv16: N too large");
    GvParams p;
    p.x = a.x;
    p.stats = a.stats;
    p.stats_parts = a.stats_parts;
    p.gamma = a.gamma;
    p.beta = a.beta;
    p.a = a.a;
    p.c1 = a.c1;
    p.xb_out = a.xb_out;
    p.W = a.W;
    p.bias = a.bias;
    p.out_f32 = a.out_f32;
    p.out_bf16 = a.out_bf16_gelu;
    p.resid = a.resid;
    p.stats_out = a.stats_out;
    p.row_skip = a.row_skip;
    p.fold_flag = a.fold_flag;
    p.M = a.M;
    p.N = a.N;
    p.K = a.K;
    const int epi = a.out_f32 ? 0 : (a.out_bf16_gelu ? 1 : 2);
    // (the residual-update form sizes its tiles for 64 rows whatever M is: the number of CTAs is the number of
    // partial row statistics the consumer adds up, gemv16_parts)
    const int nt = gv_pick_nt(a.N, a.K, a.x ? 1 : (a.c1 ? 2 : 0), epi, epi == 2 ? 4 : (a.M <= 16 ? 1 : (a.M <= 32 ? 2 : 4)));
    if (a.x) {
        if (epi == 0) return dispatch_mt<1, 0>(p, nt, stream);
        if (epi == 1) return dispatch_mt<1, 1>(p, nt, stream);
        return dispatch_mt<1, 2>(p, nt, stream);
    }
    if (a.c1) {
        if (epi == 0) return dispatch_mt<2, 0>(p, nt, stream);
        return dispatch_mt<2, 1>(p, nt, stream);
    }
    if (epi == 0) return dispatch_mt<0, 0>(p, nt, stream);
    if (epi == 1) return dispatch_mt<0, 1>(p, nt, stream);
    return dispatch_mt<0, 2>(p, nt, stream);
}

}  // namespace wsb
