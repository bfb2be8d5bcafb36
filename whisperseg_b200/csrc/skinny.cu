// K5d -- skinny (decode) linear layer for 65..256 rows in ONE launch: split-K across a thread-block cluster,
// partial accumulators reduced through distributed shared memory, LayerNorm folded into the projection.
//
// The tcgen05 split-K GEMM (gemm.cu) writes fp32 partial planes and needs a second launch to sum them (and a
// third of the time of a decode position at 240 rows is spent in those second-phase kernels).  Here the S CTAs
// that share an output tile form a cluster (S = 1, 2, 4 or 8 along K):
//   * every CTA runs the usual pipeline on its K slice (warp 0: TMA producer, weight tiles issued before the
//     programmatic-dependent-launch wait; warp 1: tcgen05.mma into TMEM; warps 2-5: epilogue);
//   * the epilogue warps copy their 128 x 128 fp32 partial tile from TMEM to (padded) shared memory, the cluster
//     synchronises, and CTA r sums rows [r * 128/S, (r+1) * 128/S) of all S partial tiles through DSMEM in a
//     fixed order (deterministic), applies the epilogue and writes the final values -- no planes, no second kernel;
//   * epilogues: (0/1) folded LayerNorm, out = rstd (acc - mean c1) + c2 -> fp32 | GELU -> bf16, with the row
//     statistics summed from the per-tile partials the producer of the residual stream left behind;
//     (2) in-place fp32 residual update x += acc + bias, plus a bf16 copy of the updated rows and this tile's
//     partial (sum, sum of squares) of every row for the next folded LayerNorm.
// Replaces nn.Linear (+ the preceding nn.LayerNorm) of HF WhisperDecoderLayer (modeling_whisper.py:417-506) for
// decode batches above the 64 rows gemv.cu handles; same folded weights as gemv.cu (weights.py: fold_layernorm).
#include "common.cuh"
#include "wsb_internal.h"

#include <algorithm>

namespace wsb {

namespace {

constexpr int kSkBM = 128, kSkBN = 128, kSkBK = 64;
constexpr int kSkStages = 4;
constexpr int kSkABytes = kSkBM * kSkBK * 2, kSkBBytes = kSkBN * kSkBK * 2;
constexpr int kSkStageBytes = kSkABytes + kSkBBytes;
constexpr int kSkPartLd = kSkBN + 1;                          // padded row of the partial tile (bank conflicts)
constexpr int kSkPartBytes = kSkBM * kSkPartLd * 4;
constexpr int kSkThreads = 192;                               // producer, MMA, 4 epilogue warps
constexpr int kSkSmem = kSkStages * kSkStageBytes + kSkPartBytes + 1024 /*align*/ + 256 /*barriers*/;

struct SkDev {
    int M, N, K;
    int mode;                      // 0: fold -> fp32, 1: fold -> GELU bf16, 2: residual
    const float* bias;             // [N]: c2 (fold) or bias (residual)
    const float* c1;               // [N] (fold)
    const float* stats;            // fold: [stats_parts][stats_ld][2]
    int stats_parts, stats_ld;
    float* out_f32;                // mode 0: [M][N]
    __nv_bfloat16* out_bf16;       // mode 1: [M][N]
    float* resid;                  // mode 2: [M][N], in place
    __nv_bfloat16* xb_out;         // mode 2, optional: [M][N]
    float* stats_out;              // mode 2, optional: [N / 128][stats_out_ld][2]
    int stats_out_ld;
    const unsigned char* row_skip; // optional [M]
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_dsmem(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

template <int S>
__global__ void __launch_bounds__(kSkThreads, 1)
skinny_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SkDev p) {
    extern __shared__ unsigned char sk_smem_raw[];
    unsigned char* smem = sk_smem_raw + ((1024u - (smem_u32(sk_smem_raw) & 1023u)) & 1023u);
    float* part = reinterpret_cast<float*>(smem + kSkStages * kSkStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSkStages * kSkStageBytes + kSkPartBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kSkStages;
    uint64_t* tmem_full = bars + 2 * kSkStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSkStages + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int rank = (S > 1) ? static_cast<int>(cluster_ctarank()) : 0;     // K split handled by this CTA
    const int nt = blockIdx.y, mt = blockIdx.z;
    const int nkb = p.K / kSkBK;
    const int kb0 = rank * nkb / S, kb1 = (rank + 1) * nkb / S;             // non-empty: the host guarantees nkb >= S

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kSkStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kSkBN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {
            // weights are constants: the first stages' B tiles are in flight before the dependency wait
            const int pre = min(kSkStages, kb1 - kb0);
            for (int i = 0; i < pre; ++i) {
                mbar_arrive_expect_tx(&full[i], kSkStageBytes);
                tma_load_2d(smem + i * kSkStageBytes + kSkABytes, &tmB, &full[i], (kb0 + i) * kSkBK, nt * kSkBN);
            }
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                unsigned char* sa = smem + stage * kSkStageBytes;
                if (kb - kb0 < pre) {
                    tma_load_2d(sa, &tmA, &full[stage], kb * kSkBK, mt * kSkBM);
                } else {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], kSkStageBytes);
                    tma_load_2d(sa, &tmA, &full[stage], kb * kSkBK, mt * kSkBM);
                    tma_load_2d(sa + kSkABytes, &tmB, &full[stage], kb * kSkBK, nt * kSkBN);
                }
                if (++stage == kSkStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kSkBM, kSkBN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * kSkStageBytes);
                const uint64_t da = umma_desc_k_sw128(sa);
                const uint64_t db = umma_desc_k_sw128(sa + kSkABytes);
#pragma unroll
                for (int k = 0; k < kSkBK / 16; ++k)
                    umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                umma_commit(&empty[stage]);
                if (kb == kb1 - 1) umma_commit(tmem_full);
                if (++stage == kSkStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ---- epilogue phase 1: this CTA's partial tile TMEM -> shared memory (thread = row)
        pdl_wait();
        const int q = warp & 3;                            // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < kSkBN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_base + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) part[row * kSkPartLd + c * 32 + i] = __uint_as_float(r[i]);
        }
        tc_fence_before();
    }
    // every CTA's partial tile is complete and visible to its peers
    if constexpr (S > 1) {
        cluster_arrive_release();
        cluster_wait_acquire();
    } else {
        __syncthreads();
    }

    if (warp >= 2) {
        // ---- epilogue phase 2: CTA `rank` finishes rows [r_lo, r_hi) of the tile; warp e takes every 4th row,
        // a lane the columns lane, lane + 32, ... (coalesced rows both in DSMEM and in global memory)
        constexpr int RS = kSkBM / S;
        const int e = warp - 2;
        const int valid_rows = min(kSkBM, p.M - mt * kSkBM);
        const int r_lo = rank * RS, r_hi = min(valid_rows, r_lo + RS);
        uint32_t peer[S];
#pragma unroll
        for (int s = 0; s < S; ++s) peer[s] = (S > 1) ? map_to_cta(smem_u32(part), static_cast<uint32_t>(s)) : smem_u32(part);
        const int n0 = nt * kSkBN;
        float cb[kSkBN / 32], cc1[kSkBN / 32];
#pragma unroll
        for (int c = 0; c < kSkBN / 32; ++c) {
            cb[c] = p.bias ? __ldg(p.bias + n0 + c * 32 + lane) : 0.0f;
            cc1[c] = (p.mode != 2) ? __ldg(p.c1 + n0 + c * 32 + lane) : 0.0f;
        }
        for (int row = r_lo + e; row < r_hi; row += 4) {
            const long long grow = static_cast<long long>(mt) * kSkBM + row;
            if (p.row_skip && p.row_skip[grow]) continue;                     // warp-uniform
            float v[kSkBN / 32];
#pragma unroll
            for (int c = 0; c < kSkBN / 32; ++c) v[c] = 0.0f;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const uint32_t a = peer[s] + static_cast<uint32_t>((row * kSkPartLd + lane) * 4);
#pragma unroll
                for (int c = 0; c < kSkBN / 32; ++c) {
                    if constexpr (S > 1) v[c] += ld_dsmem(a + c * 128);
                    else v[c] += part[row * kSkPartLd + c * 32 + lane];
                }
            }
            if (p.mode != 2) {
                float sm = 0.0f, sq = 0.0f;
                for (int pp = lane; pp < p.stats_parts; pp += 32) {
                    const float2 st = *reinterpret_cast<const float2*>(p.stats + (static_cast<long long>(pp) * p.stats_ld + grow) * 2);
                    sm += st.x;
                    sq += st.y;
                }
                sm = warp_sum(sm);
                sq = warp_sum(sq);
                const float mean = sm / p.K;
                const float rstd = rsqrtf(fmaxf(sq / p.K - mean * mean, 0.0f) + 1e-5f);
#pragma unroll
                for (int c = 0; c < kSkBN / 32; ++c) {
                    const float o = rstd * (v[c] - mean * cc1[c]) + cb[c];
                    const long long idx = grow * p.N + n0 + c * 32 + lane;
                    if (p.mode == 0) p.out_f32[idx] = o;
                    else p.out_bf16[idx] = __float2bfloat16(gelu_fast(o));
                }
            } else {
                float sm = 0.0f, sq = 0.0f;
#pragma unroll
                for (int c = 0; c < kSkBN / 32; ++c) {
                    const long long idx = grow * p.N + n0 + c * 32 + lane;
                    const float xn = p.resid[idx] + v[c] + cb[c];
                    p.resid[idx] = xn;
                    if (p.xb_out) p.xb_out[idx] = __float2bfloat16(xn);
                    sm += xn;
                    sq = fmaf(xn, xn, sq);
                }
                if (p.stats_out) {
                    sm = warp_sum(sm);
                    sq = warp_sum(sq);
                    if (lane == 0)
                        *reinterpret_cast<float2*>(p.stats_out + (static_cast<long long>(nt) * p.stats_out_ld + grow) * 2) = make_float2(sm, sq);
                }
            }
        }
    }
    // nobody leaves (and frees its shared memory) while a peer may still be reading it
    if constexpr (S > 1) {
        cluster_arrive_release();
        cluster_wait_acquire();
    } else {
        __syncthreads();
    }
    if (warp == 1) tmem_dealloc<kSkBN>(tmem_base);
}

template <int S>
int launch_skinny(const CUtensorMap& tmA, const CUtensorMap& tmB, const SkDev& p, int n_tiles, int m_tiles, cudaStream_t stream) {
    static PerDeviceOnce once;
    int dev = 0;
    if (once.need(&dev)) {
        WSB_CHECK_CUDA(cudaFuncSetAttribute(skinny_cluster_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkSmem));
        once.mark(dev);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S, n_tiles, m_tiles);
    cfg.blockDim = dim3(kSkThreads);
    cfg.dynamicSmemBytes = kSkSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (S > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = S;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (g_use_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    WSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, skinny_cluster_kernel<S>, tmA, tmB, p));
    count_launch();
    return 0;
}

}  // namespace

bool skinny_cluster_supported(int M, int N, int K) {
    return M >= 1 && M <= 256 && N % kSkBN == 0 && K % kSkBK == 0 && K >= kSkBK;
}

int skinny_cluster_splits(int M, int N, int K) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int clusters = ceil_div(M, kSkBM) * (N / kSkBN), nkb = K / kSkBK;
    int s = 1;
    for (int c : {2, 4, 8})
        if (clusters * c <= sms && nkb >= 2 * c) s = c;                      // >= 2 k-blocks per CTA
    return s;
}

int skinny_cluster_linear(const SkinnyArgs& a, cudaStream_t stream) {
    WSB_REQUIRE(skinny_cluster_supported(a.M, a.N, a.K), "skinny_cluster_linear: M <= 256, N % 128 == 0, K % 64 == 0");
    WSB_REQUIRE(a.A && a.W && a.lda % 8 == 0, "skinny_cluster_linear: operands");
    const int outs = (a.out_f32 != nullptr) + (a.out_bf16_gelu != nullptr) + (a.resid != nullptr);
    WSB_REQUIRE(outs == 1, "skinny_cluster_linear: exactly one output mode");
    WSB_REQUIRE(a.resid || (a.c1 && a.stats && a.stats_parts >= 1 && a.stats_ld >= a.M), "skinny_cluster_linear: folded LayerNorm inputs");
    WSB_REQUIRE(!a.stats_out || a.stats_out_ld >= a.M, "skinny_cluster_linear: stats_out_ld");
    SkDev p;
    p.M = a.M;
    p.N = a.N;
    p.K = a.K;
    p.mode = a.out_f32 ? 0 : (a.out_bf16_gelu ? 1 : 2);
    p.bias = a.bias;
    p.c1 = a.c1;
    p.stats = a.stats;
    p.stats_parts = a.stats_parts;
    p.stats_ld = a.stats_ld;
    p.out_f32 = a.out_f32;
    p.out_bf16 = a.out_bf16_gelu;
    p.resid = a.resid;
    p.xb_out = a.xb_out;
    p.stats_out = a.stats_out;
    p.stats_out_ld = a.stats_out_ld;
    p.row_skip = a.row_skip;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.M)};
        uint64_t strides[1] = {static_cast<uint64_t>(a.lda) * 2};
        uint32_t box[2] = {kSkBK, kSkBM};
        const int rc = make_tmap_bf16(&tmA, a.A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.N)};
        uint64_t strides[1] = {static_cast<uint64_t>(a.K) * 2};
        uint32_t box[2] = {kSkBK, kSkBN};
        const int rc = make_tmap_bf16(&tmB, a.W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int n_tiles = a.N / kSkBN, m_tiles = ceil_div(a.M, kSkBM);
    const int s = a.splits > 0 ? a.splits : skinny_cluster_splits(a.M, a.N, a.K);
    WSB_REQUIRE((s == 1 || s == 2 || s == 4 || s == 8) && a.K / kSkBK >= s, "skinny_cluster_linear: splits in {1, 2, 4, 8}, <= K / 64");
    switch (s) {
        case 1: return launch_skinny<1>(tmA, tmB, p, n_tiles, m_tiles, stream);
        case 2: return launch_skinny<2>(tmA, tmB, p, n_tiles, m_tiles, stream);
        case 4: return launch_skinny<4>(tmA, tmB, p, n_tiles, m_tiles, stream);
        default: return launch_skinny<8>(tmA, tmB, p, n_tiles, m_tiles, stream);
    }
}

}  // namespace wsb
