// K3 -- persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T )        A, W bf16 (K contiguous), fp32 accumulation
//
// Replaces every nn.Linear / conv2 the reference reaches through HF WhisperEncoder/Decoder
// (modeling_whisper.py:279-282, 376-377, 619-625) -- cuBLAS/cuDNN library calls there.
//
// Structure (one CTA per SM, 384 threads, tiles 128 x BN x 64):
//   warp 0   TMA producer: cp.async.bulk.tensor loads of the A and W tiles into a STAGES-deep ring
//            of 128B-swizzled shared-memory buffers, completion signalled on `full` mbarriers;
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage,
//            accumulating in TMEM; tcgen05.commit releases the stage (`empty`) and, after the last
//            k-block, publishes the accumulator (`tmem_full`);
//   warp 2   TMEM allocator (2 accumulator buffers of BN columns, so the epilogue of tile i overlaps
//            the main loop of tile i+1);
//   warps 4-11 epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> 32x32 transpose through a
//            padded shared-memory tile (so every global access is a coalesced 128-byte row) -> fused
//            bias / GELU / positional row-vector / fp32 residual -> global; the arg-max mode keeps a
//            row per thread and never materialises the [M,N] logits.
// Tiles are visited n-fastest so the A tile is read from HBM once and re-used out of L2 by the other
// n-tiles, while W (<= a few MB) stays L2 resident.
#include "common.cuh"
#include "wsb_internal.h"

namespace wsb {

constexpr int kBM = 128;
constexpr int kBK = 64;
// epilogue warps: a multiple of 4 (one set per TMEM lane quarter).  16 warps hide the epilogue of short-K
// tiles (K=1280: 20 k-blocks) at the price of one pipeline stage; long-K tiles keep 8 warps and 4 stages.
constexpr int kMaxEpiWarps = 16;
constexpr int kABytes = kBM * kBK * 2;

struct GemmDev {
    int M, N, K;
    int a_rows_per_batch;          // 0 = flat
    const float* bias;
    const float* bias2;
    int act;
    const float* resid;
    long long ldr;
    const float* rowvec;
    int out_mode;
    void* out;
    long long ldc;
    int rows_per_batch;
    float* argmax_val;
    int* argmax_idx;
    int m_tiles, n_tiles;
    int splits, kb_per_split;      // split-K: tile space is m_tiles x n_tiles x splits, fp32 partial planes
    long long split_stride;        // elements between partial planes of the output
    const unsigned char* row_skip; // optional [M]: rows flagged non-zero are not stored (finished decode rows)
    const int* n_tile_list;        // optional: only these n-tiles are computed (n_tiles = length of the list)
    int c_batch_pad;               // bf16 outputs with a_rows_per_batch > 0: extra rows between batches of C (row = global row + batch * pad)
    int group_m;                   // 0: tiles run n-fastest (A read from HBM once, B re-read per m-tile out of L2);
                                   // G > 0: when B is too big for L2, super-rows of G m-tiles are walked n-outer /
                                   // m-inner, so the CTAs in flight share a few B tiles and B streams once per super-row
};

// persistent tile index -> (split, m-tile, n-tile slot)
__device__ __forceinline__ void tile_coords(const GemmDev& p, int tile, int& sp, int& mt, int& nti) {
    sp = tile % p.splits;
    const int mn = tile / p.splits;
    if (p.group_m > 0) {
        const int per_group = p.group_m * p.n_tiles;
        const int g = mn / per_group, r = mn - g * per_group;
        const int gm = min(p.group_m, p.m_tiles - g * p.group_m);     // rows of the (possibly short) last group
        nti = r / gm;
        mt = g * p.group_m + (r - nti * gm);
    } else {
        mt = mn / p.n_tiles;
        nti = mn - mt * p.n_tiles;
    }
}

template <int BN, int EPIW>
struct GemmCfg {
    static constexpr int kEpiWarps = EPIW;
    static constexpr int kEpiGroups = EPIW / 4;          // column groups
    static constexpr int kThreads = 128 + 32 * EPIW;     // 4 control warps + epilogue warps
    static constexpr int kBBytes = BN * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN == 256) ? (EPIW > 8 ? 3 : 4) : (BN == 128) ? (EPIW > 8 ? 4 : 6) : (EPIW > 8 ? 6 : 8);
    static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
    static constexpr int kEpiBytes = kEpiWarps * 32 * 33 * 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// epilogue variants are compile-time: the per-row loops must be branch-free and small enough to stay in
// the instruction cache (runtime mode flags made the epilogue fetch- and branch-bound)
enum Epi : int { EPI_BF16 = 0, EPI_BF16_GELU, EPI_F32, EPI_F32_RESID, EPI_F32_GELU_ROWVEC, EPI_HEADMAJOR, EPI_ARGMAX };

template <int BN, int EPI, int EPIW>
__global__ void __launch_bounds__(128 + 32 * EPIW, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
    using Cfg = GemmCfg<BN, EPIW>;
    constexpr int kEpiWarps = Cfg::kEpiWarps, kEpiGroups = Cfg::kEpiGroups;
    constexpr int S = Cfg::kStages;
    extern __shared__ unsigned char smem_raw[];
    // align inside the shared window with pointer arithmetic (an integer round-trip would demote every
    // later access through this pointer to a generic LD/ST)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* epi_buf = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes + Cfg::kEpiBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tmem_full = bars + 2 * S;
    uint64_t* tmem_empty = bars + 2 * S + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kEpiWarps * 32);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (descriptor prefetch, barrier init, TMEM allocation) overlaps the previous kernel under
    // programmatic dependent launch; so do the first weight (B operand) tiles below -- weights are constants.
    // Every thread that touches activations or outputs calls pdl_wait() first.
    pdl_launch_dependents();

    const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
    const int k_blocks = p.K / kBK;
    const int tiles_per_batch = p.a_rows_per_batch > 0 ? (p.a_rows_per_batch + kBM - 1) / kBM : 0;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int pre = 0;                                // stages already armed, their B tile in flight
            if (static_cast<int>(blockIdx.x) < total_tiles) {
                const int tile = blockIdx.x;
                int sp, mt, nti;
                tile_coords(p, tile, sp, mt, nti);
                const int nt = p.n_tile_list ? p.n_tile_list[nti] : nti;
                const int kb0 = sp * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
                pre = min(S, kb1 - kb0);
                for (int i = 0; i < pre; ++i) {
                    mbar_arrive_expect_tx(&full[i], Cfg::kStageBytes);
                    tma_load_2d(smem + i * Cfg::kStageBytes + kABytes, &tmB, &full[i], (kb0 + i) * kBK, nt * BN);
                }
            }
            pdl_wait();
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int sp, mt, nti;
                tile_coords(p, tile, sp, mt, nti);
                const int nt = p.n_tile_list ? p.n_tile_list[nti] : nti;
                const int kb0 = sp * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
                int a_row, a_batch;
                if (tiles_per_batch > 0) {
                    a_batch = mt / tiles_per_batch;
                    a_row = (mt - a_batch * tiles_per_batch) * kBM;
                } else {
                    a_batch = 0;
                    a_row = mt * kBM;
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    unsigned char* sa = smem + stage * Cfg::kStageBytes;
                    if (pre > 0) {                      // armed before the dependency wait: only A is missing
                        --pre;
                        tma_load_3d(sa, &tmA, &full[stage], kb * kBK, a_row, a_batch);
                    } else {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
                        tma_load_3d(sa, &tmA, &full[stage], kb * kBK, a_row, a_batch);
                        tma_load_2d(sa + kABytes, &tmB, &full[stage], kb * kBK, nt * BN);
                    }
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
                const int sp = tile % p.splits;
                const int kb0 = sp * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
                const int as = local & 1;
                const uint32_t aphase = (local >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t da = umma_desc_k_sw128(sa);
                    const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[stage]);
                    if (kb == kb1 - 1) umma_commit(&tmem_full[as]);
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: 8 warps; warp (4 + e) owns TMEM lanes 32*(e%4).. (hardware restriction: a warp
        // may only touch the lane quarter warp_id % 4) and the 32-column chunks c with c % 2 == e / 4.
        pdl_wait();
        const int e = warp - 4;
        const int q = e & 3, half = e >> 2;        // half = column group index
        float* tbuf = epi_buf + e * (32 * 33);            // per-warp 32x32 transpose tile (padded)
        int local = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
            int sp, mt, nti;
            tile_coords(p, tile, sp, mt, nti);
            const int nt = p.n_tile_list ? p.n_tile_list[nti] : nti;       // nti: dense slot of the arg-max partials
            const int as = local & 1;
            const uint32_t aphase = (local >> 1) & 1;
            long long grow0;                            // global row of this warp's first lane
            long long crow_shift = 0;                   // C rows are shifted by this much (batches of C padded with extra rows)
            int nvalid;                                 // valid rows in the 32-row slab
            if (tiles_per_batch > 0) {
                const int b = mt / tiles_per_batch;
                const int brow0 = (mt - b * tiles_per_batch) * kBM + q * 32;
                nvalid = min(32, max(0, p.a_rows_per_batch - brow0));
                grow0 = static_cast<long long>(b) * p.a_rows_per_batch + brow0;
                crow_shift = static_cast<long long>(b) * p.c_batch_pad;
            } else {
                grow0 = static_cast<long long>(mt) * kBM + q * 32;
                nvalid = static_cast<int>(min(32LL, max(0LL, static_cast<long long>(p.M) - grow0)));
            }
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            if constexpr (EPI == EPI_ARGMAX) {
                // row-per-thread running arg-max over this tile's columns (first 4 epilogue warps only)
                if (half == 0) {
                    float best = -INFINITY;
                    int best_idx = nt * BN;
                    const bool valid = lane < nvalid;
#pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32(t_base + c * 32, r);
                        tmem_ld_wait();
                        const int n0 = nt * BN + c * 32;
                        if (!valid || n0 >= p.N) continue;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (n0 + i < p.N) {
                                float v = __uint_as_float(r[i]);
                                if (p.bias) v += __ldg(p.bias + n0 + i);
                                if (p.bias2) v += __ldg(p.bias2 + n0 + i);
                                if (v > best) {
                                    best = v;
                                    best_idx = n0 + i;
                                }
                            }
                        }
                    }
                    if (valid) {
                        p.argmax_val[(grow0 + lane) * p.n_tiles + nti] = best;
                        p.argmax_idx[(grow0 + lane) * p.n_tiles + nti] = best_idx;
                    }
                }
            } else {
#pragma unroll 1
                for (int c = half; c < BN / 32; c += kEpiGroups) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_base + c * 32, r);
                    tmem_ld_wait();
                    const int n0 = nt * BN + c * 32;
                    if (n0 >= p.N || nvalid == 0) continue;             // warp-uniform
                    // transpose through shared memory: lane r holds row r -> lane c holds column c, so that
                    // bias / residual / output accesses are 128-byte coalesced rows
#pragma unroll
                    for (int i = 0; i < 32; ++i) tbuf[lane * 33 + i] = __uint_as_float(r[i]);
                    __syncwarp();
                    const int col = n0 + lane;
                    const bool cvalid = col < p.N;
                    float bias_v = 0.0f;
                    if (cvalid && p.bias) bias_v = __ldg(p.bias + col);
                    if constexpr (EPI == EPI_F32_RESID) {
                        // the residual aliases the output (in-place stream update): all loads are issued ahead of
                        // the stores explicitly, the compiler may not reorder them.  lane = 8 * rs + j owns columns
                        // 4j..4j+3 of rows rs, rs + 4, ...: 16-byte accesses, 128 contiguous bytes per row
                        const int rs4 = lane >> 3, j4 = (lane & 7) * 4;
                        const int c4 = n0 + j4;
                        const bool vec4 = c4 + 4 <= p.N && (p.ldc & 3) == 0 && (p.ldr & 3) == 0 &&
                                          ((reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.resid)) & 15) == 0;
                        if (__all_sync(0xffffffffu, vec4)) {
                            float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.bias) {
                                bz.x = __ldg(p.bias + c4);
                                bz.y = __ldg(p.bias + c4 + 1);
                                bz.z = __ldg(p.bias + c4 + 2);
                                bz.w = __ldg(p.bias + c4 + 3);
                            }
                            const float* rs = p.resid + (grow0 + rs4) * p.ldr + c4;
                            float* o = reinterpret_cast<float*>(p.out) + (grow0 + rs4) * p.ldc + c4;
                            float4 res[8];
#pragma unroll
                            for (int it = 0; it < 8; ++it)
                                res[it] = (it * 4 + rs4 < nvalid) ? *reinterpret_cast<const float4*>(rs + static_cast<long long>(it * 4) * p.ldr)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int rr = it * 4 + rs4;
                                if (rr < nvalid) {
                                    float4 v;
                                    v.x = tbuf[rr * 33 + j4] + bz.x + res[it].x;
                                    v.y = tbuf[rr * 33 + j4 + 1] + bz.y + res[it].y;
                                    v.z = tbuf[rr * 33 + j4 + 2] + bz.z + res[it].z;
                                    v.w = tbuf[rr * 33 + j4 + 3] + bz.w + res[it].w;
                                    *reinterpret_cast<float4*>(o + static_cast<long long>(it * 4) * p.ldc) = v;
                                }
                            }
                        } else {
                            const float* rs = p.resid + grow0 * p.ldr + col;
                            float* o = reinterpret_cast<float*>(p.out) + grow0 * p.ldc + col;
#pragma unroll
                            for (int r0 = 0; r0 < 32; r0 += 16) {
                                float res[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    res[j] = (cvalid && r0 + j < nvalid) ? rs[static_cast<long long>(r0 + j) * p.ldr] : 0.0f;
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (cvalid && r0 + j < nvalid)
                                        o[static_cast<long long>(r0 + j) * p.ldc] = tbuf[(r0 + j) * 33 + lane] + bias_v + res[j];
                            }
                        }
                    } else if constexpr (EPI == EPI_F32 || EPI == EPI_F32_GELU_ROWVEC) {
                        float* o = reinterpret_cast<float*>(p.out) + sp * p.split_stride + grow0 * p.ldc + col;
                        // finished decode rows: skip the partial-plane stores (lane rr holds row rr's flag)
                        unsigned skip_mask = 0;
                        if constexpr (EPI == EPI_F32) {
                            if (p.row_skip)
                                skip_mask = __ballot_sync(0xffffffffu, lane < nvalid && p.row_skip[grow0 + lane] != 0);
                        }
                        int brow = 0;
                        if constexpr (EPI == EPI_F32_GELU_ROWVEC) brow = static_cast<int>(grow0 % p.rows_per_batch);
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) {
                            if (rr >= nvalid) break;
                            float v = tbuf[rr * 33 + lane] + bias_v;
                            if constexpr (EPI == EPI_F32_GELU_ROWVEC) {
                                v = gelu_fast(v);
                                if (cvalid) v += __ldg(p.rowvec + static_cast<long long>(brow) * p.N + col);
                                if (++brow == p.rows_per_batch) brow = 0;
                            }
                            if (cvalid && !((skip_mask >> rr) & 1u)) o[static_cast<long long>(rr) * p.ldc] = v;
                        }
                    } else {   // bf16 outputs: EPI_BF16, EPI_BF16_GELU, EPI_HEADMAJOR
                        // lane = 4 * rsub + j owns columns 8j..8j+7 of rows rsub, rsub + 8, ...: one 16-byte store per
                        // lane, 64 contiguous bytes per row and instruction, 4 store instructions per 32x32 chunk
                        // (tbuf reads are conflict-free: bank = (row + 8j + i) mod 32 over the 32 lanes)
                        const int rsub = lane >> 2, j8 = (lane & 3) * 8;
                        const int c8 = n0 + j8;
                        bool vec = c8 + 8 <= p.N && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
                        if constexpr (EPI != EPI_HEADMAJOR) vec = vec && (p.ldc & 7) == 0;
                        float b8[8];
                        if (p.bias && c8 + 8 <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
                            const float4 u0 = __ldg(reinterpret_cast<const float4*>(p.bias + c8));
                            const float4 u1 = __ldg(reinterpret_cast<const float4*>(p.bias + c8 + 4));
                            b8[0] = u0.x; b8[1] = u0.y; b8[2] = u0.z; b8[3] = u0.w;
                            b8[4] = u1.x; b8[5] = u1.y; b8[6] = u1.z; b8[7] = u1.w;
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) b8[i] = (p.bias && c8 + i < p.N) ? __ldg(p.bias + c8 + i) : 0.0f;
                        }
                        long long hm_b = 0;
                        int hm_t = 0;
                        if constexpr (EPI == EPI_HEADMAJOR) {
                            hm_b = grow0 / p.rows_per_batch;
                            hm_t = static_cast<int>(grow0 - hm_b * p.rows_per_batch);
                        }
                        auto store_rows = [&](int it) {
                            const int rr = it * 8 + rsub;
                            if (rr >= nvalid) return;
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                v[i] = tbuf[rr * 33 + j8 + i] + b8[i];
                                if constexpr (EPI == EPI_BF16_GELU) v[i] = gelu_fast(v[i]);
                            }
                            __nv_bfloat16* o;
                            if constexpr (EPI == EPI_HEADMAJOR) {
                                int t = hm_t + rr;
                                long long b = hm_b;
                                if (t >= p.rows_per_batch) {           // a 32-row slab crosses at most one batch boundary
                                    t -= p.rows_per_batch;
                                    b += 1;
                                }
                                o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                    ((b * (p.N >> 6) + (c8 >> 6)) * p.rows_per_batch + t) * 64 + (c8 & 63);
                            } else {
                                o = reinterpret_cast<__nv_bfloat16*>(p.out) + (grow0 + crow_shift + rr) * p.ldc + c8;
                            }
                            if (vec) {
                                uint4 w;
                                w.x = pack_bf16x2(v[0], v[1]);
                                w.y = pack_bf16x2(v[2], v[3]);
                                w.z = pack_bf16x2(v[4], v[5]);
                                w.w = pack_bf16x2(v[6], v[7]);
                                *reinterpret_cast<uint4*>(o) = w;
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (c8 + i < p.N) o[i] = __float2bfloat16(v[i]);
                            }
                        };
#pragma unroll
                        for (int it = 0; it < 4; ++it) store_rows(it);
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    WSB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    WSB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        WSB_REQUIRE((gstr[i] & 15) == 0, "TMA strides must be multiples of 16 bytes");
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                    gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
        return 4;
    }
    return 0;
}

int gemm_pick_block_n(int M, int N) {
    // big GEMMs: widest tile (least operand traffic per flop); skinny (decode) GEMMs: enough tiles to
    // keep most SMs streaming weights.  Tiles that would be mostly padding along N are avoided.
    const int m_tiles = ceil_div(M, kBM);
    int best = 32, best_tiles = -1;
    for (int bn : {256, 128, 64, 32}) {
        const int nt = ceil_div(N, bn);
        if (static_cast<double>(nt) * bn > 1.15 * N && bn != 32) continue;
        const int tiles = m_tiles * nt;
        if (tiles >= 120) return bn;
        if (tiles > best_tiles) {
            best_tiles = tiles;
            best = bn;
        }
    }
    return best;
}
int gemm_n_tiles(int N, int block_n) { return ceil_div(N, block_n); }

template <int BN, int EPI, int EPIW>
static int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, EPIW>;
    static PerDeviceOnce once;
    static int num_sms = 0;
    int dev = 0;
    if (once.need(&dev)) {
        WSB_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI, EPIW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        WSB_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        once.mark(dev);
    }
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[3], strides[2];
        uint32_t box[3] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(kBM), 1};
        dims[0] = static_cast<uint64_t>(a.K);
        if (a.a_rows_per_batch > 0) {
            dims[1] = static_cast<uint64_t>(a.a_rows_per_batch);
            dims[2] = static_cast<uint64_t>(a.M / a.a_rows_per_batch);
            strides[1] = static_cast<uint64_t>(a.a_batch_stride) * 2;
        } else {
            dims[1] = static_cast<uint64_t>(a.M);
            dims[2] = 1;
            strides[1] = static_cast<uint64_t>(a.M) * static_cast<uint64_t>(a.lda) * 2;
        }
        strides[0] = static_cast<uint64_t>(a.lda) * 2;
        int rc = make_tmap_bf16(&tmA, a.A, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.N)};
        uint64_t strides[1] = {static_cast<uint64_t>(a.K) * 2};
        uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(BN)};
        int rc = make_tmap_bf16(&tmB, a.W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    GemmDev p;
    p.M = a.M;
    p.N = a.N;
    p.K = a.K;
    p.a_rows_per_batch = a.a_rows_per_batch;
    p.bias = a.bias;
    p.bias2 = a.bias2;
    p.act = a.act;
    p.resid = a.resid;
    p.ldr = a.ldr;
    p.rowvec = a.rowvec;
    p.out_mode = a.out_mode;
    p.out = a.out;
    p.ldc = a.ldc;
    p.rows_per_batch = a.rows_per_batch;
    p.argmax_val = a.argmax_val;
    p.argmax_idx = a.argmax_idx;
    if (a.a_rows_per_batch > 0)
        p.m_tiles = (a.M / a.a_rows_per_batch) * ceil_div(a.a_rows_per_batch, kBM);
    else
        p.m_tiles = ceil_div(a.M, kBM);
    p.n_tiles = a.n_tile_list ? a.n_tile_count : ceil_div(a.N, BN);
    p.n_tile_list = a.n_tile_list;
    p.splits = a.splits > 1 ? a.splits : 1;
    p.kb_per_split = ceil_div(a.K / kBK, p.splits);
    p.splits = ceil_div(a.K / kBK, p.kb_per_split);          // drop empty trailing splits
    p.split_stride = a.split_stride;
    p.row_skip = a.row_skip;
    p.c_batch_pad = a.c_batch_pad;
    // B (weights) beyond half of the 126 MB L2 (the cross-K/V projection of all decoder layers: 210 MB): walk
    // super-rows of 16 m-tiles so that B streams from HBM once per super-row instead of once per m-tile
    p.group_m = (static_cast<double>(a.N) * a.K * 2.0 > 64e6 && p.m_tiles > 1) ? 16 : 0;
    const int total = p.m_tiles * p.n_tiles * p.splits;
    const int grid = std::min(total, std::max(8, num_sms - g_sm_reserve));
    WSB_CHECK_CUDA(launch_kernel(gemm_kernel<BN, EPI, EPIW>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, tmA, tmB, p));
    count_launch();
    return 0;
}

int gemm_bf16(const GemmArgs& a, cudaStream_t stream) {
    WSB_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "empty GEMM");
    WSB_REQUIRE(a.K % kBK == 0, "K must be a multiple of 64");
    WSB_REQUIRE(a.lda % 8 == 0, "lda must be a multiple of 8 elements (16 bytes)");
    WSB_REQUIRE(a.a_rows_per_batch == 0 || a.M % a.a_rows_per_batch == 0, "M must be batches * rows_per_batch");
    if (a.out_mode == GEMM_OUT_HEADMAJOR)
        WSB_REQUIRE(a.N % 64 == 0 && a.rows_per_batch > 0 && a.M % a.rows_per_batch == 0, "head-major output shape");
    if (a.resid || a.rowvec) WSB_REQUIRE(a.N % 32 == 0, "residual / row-vector epilogues need N % 32 == 0");
    if (a.out_mode == GEMM_OUT_F32) WSB_REQUIRE(a.ldc % 4 == 0, "ldc must be a multiple of 4 for fp32 output");
    if (a.out_mode == GEMM_OUT_BF16) WSB_REQUIRE(a.ldc % 8 == 0, "ldc must be a multiple of 8 for bf16 output");
    int bn = a.block_n ? a.block_n : gemm_pick_block_n(a.M, a.N);
    // map the requested epilogue onto a compiled variant
    int epi = -1;
    if (a.out_mode == GEMM_OUT_ARGMAX && !a.resid && !a.rowvec && a.act == GEMM_ACT_NONE) epi = EPI_ARGMAX;
    else if (a.bias2) epi = -1;
    else if (a.out_mode == GEMM_OUT_HEADMAJOR && !a.resid && !a.rowvec && a.act == GEMM_ACT_NONE) epi = EPI_HEADMAJOR;
    else if (a.out_mode == GEMM_OUT_BF16 && !a.resid && !a.rowvec) epi = a.act == GEMM_ACT_GELU ? EPI_BF16_GELU : EPI_BF16;
    else if (a.out_mode == GEMM_OUT_F32 && a.resid && !a.rowvec && a.act == GEMM_ACT_NONE) epi = EPI_F32_RESID;
    else if (a.out_mode == GEMM_OUT_F32 && !a.resid && a.rowvec && a.act == GEMM_ACT_GELU) epi = EPI_F32_GELU_ROWVEC;
    else if (a.out_mode == GEMM_OUT_F32 && !a.resid && !a.rowvec && a.act == GEMM_ACT_NONE) epi = EPI_F32;
    if (epi < 0) {
        set_last_error("gemm_bf16: unsupported epilogue combination");
        return 2;
    }
    if (a.rowvec) WSB_REQUIRE(a.rows_per_batch > 0, "row-vector epilogue needs rows_per_batch");
    if (a.splits > 1) WSB_REQUIRE(epi == EPI_F32 && !a.bias && a.split_stride >= static_cast<int64_t>(a.M) * a.ldc,
                                  "split-K writes raw fp32 partial planes (no bias / activation)");
    if (a.rows_per_batch > 0) WSB_REQUIRE(a.rows_per_batch >= 32, "rows_per_batch must be >= 32");
    // short-K tiles are epilogue-bound: give them 16 epilogue warps; long-K and skinny GEMMs keep 8
    const bool wide_epi = (a.K / kBK) / (a.splits > 1 ? a.splits : 1) <= 40 && a.M > 512 && epi != EPI_ARGMAX;
#define WSB_GEMM_EPI(BN_, EPI_) (wide_epi ? launch_gemm<BN_, EPI_, 16>(a, stream) : launch_gemm<BN_, EPI_, 8>(a, stream))
#define WSB_GEMM_CASE(BN_)                                                           \
    case BN_:                                                                        \
        switch (epi) {                                                               \
            case EPI_BF16: return WSB_GEMM_EPI(BN_, EPI_BF16);                       \
            case EPI_BF16_GELU: return WSB_GEMM_EPI(BN_, EPI_BF16_GELU);             \
            case EPI_F32: return WSB_GEMM_EPI(BN_, EPI_F32);                         \
            case EPI_F32_RESID: return WSB_GEMM_EPI(BN_, EPI_F32_RESID);             \
            case EPI_F32_GELU_ROWVEC: return WSB_GEMM_EPI(BN_, EPI_F32_GELU_ROWVEC); \
            case EPI_HEADMAJOR: return WSB_GEMM_EPI(BN_, EPI_HEADMAJOR);             \
            default: return launch_gemm<BN_, EPI_ARGMAX, 8>(a, stream);              \
        }
    switch (bn) {
        WSB_GEMM_CASE(256)
        WSB_GEMM_CASE(128)
        WSB_GEMM_CASE(64)
        WSB_GEMM_CASE(32)
        default: set_last_error("unsupported block_n"); return 2;
    }
#undef WSB_GEMM_EPI
#undef WSB_GEMM_CASE
}

}  // namespace wsb

namespace wsb {

int gemm_effective_splits(int K, int splits) {
    const int kb = K / kBK;
    const int s = splits > 1 ? splits : 1;
    const int per = ceil_div(kb, s);
    return ceil_div(kb, per);
}

void gemm_pick_skinny(int M, int N, int K, int* block_n, int* splits) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int m_tiles = ceil_div(M, kBM), kb = K / kBK;
    int best_bn = 128, best_s = 1, best_ctas = 0;
    for (int bn : {128, 256}) {
        if (N % bn != 0) continue;
        const int base = m_tiles * (N / bn);
        for (int s = 1; s <= kb && s <= 16; ++s) {
            const int eff = gemm_effective_splits(K, s);
            const int ctas = base * eff;
            if (ctas > sms) break;
            if (ctas > best_ctas) {
                best_ctas = ctas;
                best_bn = bn;
                best_s = eff;
            }
        }
    }
    *block_n = best_bn;
    *splits = best_s;
}

}  // namespace wsb
