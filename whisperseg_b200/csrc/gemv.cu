// K5c -- skinny linear layers for decode batches of at most 16 rows (the tail of a greedy decode after batch
// compaction, or a short recording with a handful of windows).
//
// At <= 16 rows the tcgen05 split-K GEMM + second-phase reduce pair is pure fixed latency (TMEM allocation,
// tensor-map fetch, pipeline fill, plane write + re-read, two launches): ~13 us per linear layer against
// ~2 us of weight streaming.  This kernel does the whole layer in one launch of <= 148 CTAs (one per SM):
//   * a CTA owns 8*NT output features and the full K extent -> no split-K, no partial planes.  Its weight
//     tile is fetched by TMA bulk copies (one per weight row, padded destination rows so that the fragment
//     reads are bank-conflict free), all issued at kernel entry on one mbarrier: maximum memory-level
//     parallelism at zero register cost;
//   * the consumer's LayerNorm is fused in WITHOUT a separate pass: the producer of the residual stream (the
//     previous residual-update launch of this kernel, or the embedding kernel) leaves per-CTA partial
//     (sum, sum of squares) of every row; the consumer adds them in a fixed order, and normalises exactly the
//     fp32 elements each lane needs for its MMA fragments, while the weights are still in flight;
//   * 16 rows are exactly the M of a warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate): the 16 warps
//     split K (the k order inside a 32-element block is permuted identically for both operands so that
//     every lane's 8 contiguous elements feed two MMAs), the 16 partial tiles are summed through shared
//     memory and the epilogue (bias; GELU -> bf16 | fp32 | in-place fp32 residual add + row statistics for
//     the next LayerNorm) is applied.
// HBM-bound by design: it streams each weight byte once; the decode step is launch-latency-bound around it.
// (tcgen05 needs M = 128 tiles and a TMEM prologue -- the wrong tool at M <= 16; this is deliberate.)
#include "common.cuh"
#include "wsb_internal.h"

#include <algorithm>

namespace wsb {

constexpr int kGvThreads = 512;
constexpr int kGvWarps = kGvThreads / 32;
constexpr int kGvBatch = 3;             // k-blocks (of 32) per warp whose activation fragments are in flight
constexpr int kGvMaxNT = 5;
constexpr int kGvWPad = 64;             // bytes of padding per weight row in smem (row shift = 16 banks)

struct GvParams {
    const float* x;                     // LN input: fp32 [M][K] (IN_LN) ...
    const float* stats;                 // [parts][16][2] partial (sum, sum of squares) of every row of x
    int stats_parts;
    const float* gamma;
    const float* beta;
    const __nv_bfloat16* a;             // ... or bf16 activations [M][K]
    const __nv_bfloat16* W;             // [N][K]
    const float* bias;                  // [N] or null
    float* out_f32;                     // EPI 0: [M][N]
    __nv_bfloat16* out_bf16;            // EPI 1: [M][N] = gelu(.)
    float* resid;                       // EPI 2: [M][N] += .
    float* stats_out;                   // EPI 2: [gridDim.x][16][2] partial row statistics of the updated rows
    const unsigned char* row_skip;      // [M] or null: rows not stored
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

template <int IN_LN, int EPI, int NT>
__global__ void __launch_bounds__(kGvThreads, 1) gemv16_kernel(const GvParams p) {
    extern __shared__ __align__(128) unsigned char gv_smem[];
    __shared__ float red[kGvWarps][NT][16][8];
    __shared__ float s_new[EPI == 2 ? NT : 1][16][8];
    __shared__ float s_mean[16], s_rstd[16];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * 8 * NT;
    const int nkb = p.K >> 5;
    const int w_stride = p.K * 2 + kGvWPad;              // bytes
    unsigned char* w_s = gv_smem;
    float* gam_s = reinterpret_cast<float*>(gv_smem + static_cast<size_t>(8 * NT) * w_stride);
    float* bet_s = gam_s + p.K;
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
    uint4 a_lo[kGvBatch], a_hi[kGvBatch];
    if constexpr (IN_LN) {
        for (int j = tid; j < (p.K >> 2); j += kGvThreads) {
            reinterpret_cast<float4*>(gam_s)[j] = __ldg(reinterpret_cast<const float4*>(p.gamma) + j);
            reinterpret_cast<float4*>(bet_s)[j] = __ldg(reinterpret_cast<const float4*>(p.beta) + j);
        }
    }
    pdl_wait();

    if constexpr (IN_LN) {
        // fp32 fragments of this warp's k-blocks (K <= 1536 -> at most 3 per warp), normalised on the fly
        float4 xl[kGvBatch][2], xh[kGvBatch][2];
#pragma unroll
        for (int i = 0; i < kGvBatch; ++i) {
            const int kb = warp + i * kGvWarps;
            xl[i][0] = xl[i][1] = xh[i][0] = xh[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kb < nkb) {
                const int ko = kb * 32 + t * 8;
                if (g < p.M) {
                    const float4* s = reinterpret_cast<const float4*>(p.x + static_cast<long long>(g) * p.K + ko);
                    xl[i][0] = s[0];
                    xl[i][1] = s[1];
                }
                if (g + 8 < p.M) {
                    const float4* s = reinterpret_cast<const float4*>(p.x + static_cast<long long>(g + 8) * p.K + ko);
                    xh[i][0] = s[0];
                    xh[i][1] = s[1];
                }
            }
        }
        {   // row `warp`: add the producer's partial statistics in a fixed order
            float sm = 0.0f, sq = 0.0f;
            for (int pp = lane; pp < p.stats_parts; pp += 32) {
                const float2 v = *reinterpret_cast<const float2*>(p.stats + (static_cast<long long>(pp) * 16 + warp) * 2);
                sm += v.x;
                sq += v.y;
            }
            sm = warp_sum(sm);
            sq = warp_sum(sq);
            if (lane == 0) {
                const float mean = sm / p.K;
                const float var = fmaxf(sq / p.K - mean * mean, 0.0f);
                s_mean[warp] = mean;
                s_rstd[warp] = rsqrtf(var + 1e-5f);
            }
        }
        __syncthreads();
        const float m_lo = s_mean[g], r_lo = s_rstd[g], m_hi = s_mean[g + 8], r_hi = s_rstd[g + 8];
#pragma unroll
        for (int i = 0; i < kGvBatch; ++i) {
            const int kb = warp + i * kGvWarps;
            a_lo[i] = zero4;
            a_hi[i] = zero4;
            if (kb < nkb) {
                const int ko = kb * 32 + t * 8;
                if (g < p.M) a_lo[i] = ln_pack8(xl[i][0], xl[i][1], m_lo, r_lo, gam_s + ko, bet_s + ko);
                if (g + 8 < p.M) a_hi[i] = ln_pack8(xh[i][0], xh[i][1], m_hi, r_hi, gam_s + ko, bet_s + ko);
            }
        }
    }

    bool w_ready = false;
    for (int kb0 = warp; kb0 < nkb; kb0 += kGvBatch * kGvWarps) {
        if constexpr (!IN_LN) {
#pragma unroll
            for (int i = 0; i < kGvBatch; ++i) {
                const int kb = kb0 + i * kGvWarps;
                a_lo[i] = zero4;
                a_hi[i] = zero4;
                if (kb < nkb) {
                    const int ko = kb * 32 + t * 8;
                    if (g < p.M) a_lo[i] = *reinterpret_cast<const uint4*>(p.a + static_cast<long long>(g) * p.K + ko);
                    if (g + 8 < p.M) a_hi[i] = *reinterpret_cast<const uint4*>(p.a + static_cast<long long>(g + 8) * p.K + ko);
                }
            }
        }
        if (!w_ready) {
            mbar_wait(&bar, 0);
            w_ready = true;
        }
#pragma unroll
        for (int i = 0; i < kGvBatch; ++i) {
            const int kb = kb0 + i * kGvWarps;
            if (kb < nkb) {
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
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        red[warp][nt][g][2 * t] = acc[nt][0];
        red[warp][nt][g][2 * t + 1] = acc[nt][1];
        red[warp][nt][g + 8][2 * t] = acc[nt][2];
        red[warp][nt][g + 8][2 * t + 1] = acc[nt][3];
    }
    __syncthreads();
    for (int idx = tid; idx < NT * 128; idx += kGvThreads) {
        const int nt = idx >> 7, r = (idx & 127) >> 3, c = idx & 7;
        const int n = n0 + nt * 8 + c;
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kGvWarps; ++w) v += red[w][nt][r][c];
        const bool valid = r < p.M && n < p.N;
        if (valid && p.bias) v += __ldg(p.bias + n);
        const bool store = valid && !(p.row_skip && p.row_skip[r]);
        const long long o = static_cast<long long>(r) * p.N + n;
        if constexpr (EPI == 0) {
            if (store) p.out_f32[o] = v;
        }
        if constexpr (EPI == 1) {
            if (store) p.out_bf16[o] = __float2bfloat16(gelu_fast(v));
        }
        if constexpr (EPI == 2) {
            float xn = 0.0f;
            if (valid) {
                xn = p.resid[o] + (store ? v : 0.0f);
                if (store) p.resid[o] = xn;
            }
            s_new[nt][r][c] = xn;
        }
    }
    if constexpr (EPI == 2) {
        if (p.stats_out) {
            __syncthreads();
            if (tid < 16) {
                float sm = 0.0f, sq = 0.0f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float v = s_new[nt][tid][c];
                        sm += v;
                        sq = fmaf(v, v, sq);
                    }
                *reinterpret_cast<float2*>(p.stats_out + (static_cast<long long>(blockIdx.x) * 16 + tid) * 2) = make_float2(sm, sq);
            }
        }
    }
}

// exact (sum, sum of squares) of every row: the statistics of a residual stream that no gemv16 launch produced
__global__ void row_stats_kernel(const float* __restrict__ x, int K, float* __restrict__ stats) {
    const int r = blockIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    float sm = 0.0f, sq = 0.0f;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const float v = x[static_cast<long long>(r) * K + j];
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

int row_stats16(const float* x, int M, int K, float* stats, cudaStream_t stream) {
    WSB_REQUIRE(M >= 1 && M <= 16, "row_stats16 handles at most 16 rows");
    WSB_CHECK_CUDA(launch_kernel(row_stats_kernel, dim3(M), dim3(256), 0, stream, x, K, stats));
    count_launch();
    return 0;
}

int gemv16_parts(int N) {                                // CTAs (= partial statistics) of a launch with N outputs
    const int nt = std::min(kGvMaxNT, std::max(1, ceil_div(ceil_div(N, 8), 148)));
    return ceil_div(N, 8 * nt);
}

template <int IN_LN, int EPI, int NT>
static int launch_gemv(const GvParams& p, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(8 * NT) * (static_cast<size_t>(p.K) * 2 + kGvWPad) + (IN_LN ? static_cast<size_t>(p.K) * 8 : 0);
    WSB_REQUIRE(smem <= 180 * 1024, "gemv16: weight tile does not fit in shared memory (K too large)");
    static PerDeviceOnce once;
    int dev = 0;
    if (once.need(&dev)) {
        WSB_CHECK_CUDA(cudaFuncSetAttribute(gemv16_kernel<IN_LN, EPI, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
        once.mark(dev);
    }
    WSB_CHECK_CUDA(launch_kernel(gemv16_kernel<IN_LN, EPI, NT>, dim3(ceil_div(p.N, 8 * NT)), dim3(kGvThreads), smem, stream, p));
    count_launch();
    return 0;
}

template <int IN_LN, int EPI>
static int dispatch_nt(const GvParams& p, int nt, cudaStream_t stream) {
    switch (nt) {
        case 1: return launch_gemv<IN_LN, EPI, 1>(p, stream);
        case 2: return launch_gemv<IN_LN, EPI, 2>(p, stream);
        case 3: return launch_gemv<IN_LN, EPI, 3>(p, stream);
        case 4: return launch_gemv<IN_LN, EPI, 4>(p, stream);
        default: return launch_gemv<IN_LN, EPI, 5>(p, stream);
    }
}

int gemv16(const Gemv16Args& a, cudaStream_t stream) {
    WSB_REQUIRE(a.M >= 1 && a.M <= 16, "gemv16 handles at most 16 rows");
    WSB_REQUIRE(a.K % 32 == 0 && a.K >= 32, "gemv16: K must be a multiple of 32");
    WSB_REQUIRE((a.x != nullptr) != (a.a != nullptr), "gemv16: exactly one of the fp32 (LayerNorm) and bf16 inputs");
    WSB_REQUIRE(!a.x || (a.K <= 1536 && a.gamma && a.beta && a.stats && a.stats_parts >= 1),
                "gemv16: fused LayerNorm needs gamma/beta, row statistics and K <= 1536");
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
    p.W = a.W;
    p.bias = a.bias;
    p.out_f32 = a.out_f32;
    p.out_bf16 = a.out_bf16_gelu;
    p.resid = a.resid;
    p.stats_out = a.stats_out;
    p.row_skip = a.row_skip;
    p.M = a.M;
    p.N = a.N;
    p.K = a.K;
    const int nt = std::min(kGvMaxNT, std::max(1, ceil_div(ceil_div(a.N, 8), 148)));
    const int epi = a.out_f32 ? 0 : (a.out_bf16_gelu ? 1 : 2);
    if (a.x) {
        if (epi == 0) return dispatch_nt<1, 0>(p, nt, stream);
        if (epi == 1) return dispatch_nt<1, 1>(p, nt, stream);
        return dispatch_nt<1, 2>(p, nt, stream);
    }
    if (epi == 0) return dispatch_nt<0, 0>(p, nt, stream);
    if (epi == 1) return dispatch_nt<0, 1>(p, nt, stream);
    return dispatch_nt<0, 2>(p, nt, stream);
}

}  // namespace wsb
