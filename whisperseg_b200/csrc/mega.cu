// K5e -- one persistent kernel per decoder position for decode batches of at most 64 rows.
//
// At <= 64 rows a decoder position on the fused path (gemv.cu + decode.cu) is 8 dependent launches per layer --
// 258 launches of 6-8 us each for ~0.3 ms of HBM traffic (whisper-large): the position is launch-latency-bound.
// This kernel runs the whole layer stack of one position in ONE launch of one 512-thread CTA per SM:
//   phase 0      token + position embedding -> fp32 residual stream, its bf16 copy and exact row statistics
//   per layer    qkv (folded LayerNorm) | self-attention (+ K/V cache append) | out-proj (+ residual, bf16 copy,
//                row statistics) | cross-q (folded LayerNorm) | cross-attention | cross-out | fc1 (folded
//                LayerNorm, GELU) | fc2 (+ residual ...)                         -- 8 phases
// with a grid barrier between phases (one monotonic counter in global memory, ld.acquire spin, watchdog).
// What the launches could not do and the persistent kernel can: the weight tile of a CTA's NEXT TWO linear-layer
// jobs is already in flight (TMA bulk copies into a two-slot shared-memory ring, armed before griddepcontrol.wait
// for the first two) while the current phase computes, so a phase only waits for the barrier, its activations
// (L2 hits, ld.global.cg -- the L1 is not coherent across CTAs) and its own MMAs.
// The arithmetic is gemv.cu's and decode.cu's, operation for operation (same k split over the 16 / MT warps, same
// fixed-order reductions, per-8-feature row-statistics partials), so the tokens are bit-identical to the launch-per-
// layer path: tests/test_gpu_mega.py.  Replaces, per generated token, HF WhisperDecoderLayer x n_layers
// (modeling_whisper.py:417-506) inside the generate loop (reference model.py:655-666).
// All CTAs of the grid must be resident at once (grid = #SMs, 1 CTA/SM): two of these kernels on one device at the
// same time could starve each other, so the engine serialises them per device (engine.cu).
#include "common.cuh"
#include "wsb_internal.h"
#include "decode.h"

#include <algorithm>

namespace wsb {

constexpr int kMgThreads = 512;
constexpr int kMgWarps = kMgThreads / 32;
constexpr int kMgBatch = 3;                 // k-blocks (of 32) per warp whose activation fragments are in flight
constexpr int kMgMaxNT = 4;
constexpr int kMgWPad = 64;                 // bytes of padding per weight row in smem (as gemv.cu)
constexpr int kMgSlotBytes = 88 * 1024;     // one weight-ring slot
constexpr int kMgGroups = kMgThreads / 128; // attention units processed concurrently per CTA
constexpr int kMgMaxKeys = 512;

struct MegaOp {                             // one linear layer of a decoder layer
    const __nv_bfloat16* W;                 // [N][K]  (LayerNorm-consuming ops: bf16(W o gamma))
    const float* bias;                      // [N]     (LayerNorm-consuming ops: c2 = b + W beta)
    const float* c1;                        // [N] row sums of W (LayerNorm-consuming ops), else null
};
struct MegaLayerDev {
    MegaOp op[6];                           // qkv, self-out, cross-q, cross-out, fc1, fc2
};

struct MegaParams {
    const MegaLayerDev* layers;
    int L, d, F, H, T, tmax, B;
    const int* next_token;
    const int* step_ptr;
    const __nv_bfloat16* emb;
    const float* pos_emb;
    float* dx;                              // [B][d] fp32 residual stream
    __nv_bfloat16* dxn;                     // [B][d] its bf16 copy (input of the folded-LayerNorm projections)
    float* stats;                           // [parts][MP][2] partial (sum, sum of squares) of every row of dx
    float* proj;                            // [B][3d] fp32 projection outputs (qkv / cross-q)
    __nv_bfloat16* datt;                    // [B][d] attention output
    __nv_bfloat16* dff;                     // [B][F]
    __nv_bfloat16* k_cache;                 // [L][B][H][tmax][64]
    __nv_bfloat16* v_cache;
    const __nv_bfloat16* cross_kv;          // [B / kv_div][L][2][H][T][64]
    const unsigned char* finished;          // [B] or null
    int kv_div;
    unsigned int* sync;                     // [0] barrier arrivals, [1] exits, [2] watchdog flag, [32] last completed barrier
    int* fold_flag;                         // folded-LayerNorm guard (see gemv.cu) or null
    int nt[6];                              // n-tiles (of 8 features) per CTA job, per op
    unsigned long long* trace;              // diagnostics (or null): CTA 0 stamps %globaltimer after every grid barrier
    int stage_base;                         // ... and accumulates per-stage times of its linear-layer jobs from this word on
};

__device__ __forceinline__ void mg_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mg_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned int mg_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mg_group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); }

// Grid barrier.  Arrivals are atomic adds on sync[0]; the LAST arriver of barrier k publishes k in sync[32] (another
// 128-byte line) and everybody else polls that word with a short sleep between polls -- 147 CTAs hammering the line the
// atomics go to would serialise the arrivals behind the polling traffic (measured: 18 us per phase that way).
// A CTA that waits longer than ~2 s raises the watchdog flag and all CTAs run to the end without waiting any more
// (the host turns the flag into an error).
__device__ __forceinline__ unsigned long long mg_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mg_grid_sync(unsigned int* sync, unsigned int k, unsigned long long* trace = nullptr) {
    if (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) trace[2 * k - 1] = mg_globaltimer();     // work of the phase done
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int old = atomicAdd(sync, 1u);
        if (old + 1u == k * gridDim.x) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sync + 32), "r"(k) : "memory");
        } else {
            const long long t0 = clock64();
            unsigned int spins = 0;
            while (mg_ld_acquire(sync + 32) < k) {
                __nanosleep(32);
                if ((++spins & 1023u) == 0u) {
                    if (mg_ld_acquire(sync + 2) != 0u) break;
                    if (clock64() - t0 > 4000000000LL) {
                        atomicExch(sync + 2, 1u);
                        break;
                    }
                }
            }
        }
        __threadfence();
        if (trace != nullptr && blockIdx.x == 0) trace[2 * k] = mg_globaltimer();                         // barrier passed
    }
    __syncthreads();
}

// the CTA's linear-layer jobs in execution order: (layer, op, tile) with tile = blockIdx.x + k * gridDim.x
struct MgCursor {
    int l, op, tile;
};
__device__ __forceinline__ int mg_op_n(const MegaParams& p, int op) { return op == 0 ? 3 * p.d : (op == 4 ? p.F : p.d); }
__device__ __forceinline__ int mg_op_k(const MegaParams& p, int op) { return op == 5 ? p.F : p.d; }
__device__ __forceinline__ bool mg_advance(const MegaParams& p, MgCursor& c) {      // to the next job; false at the end
    for (;;) {
        if (c.l >= p.L) return false;
        c.tile += gridDim.x;
        if (c.tile * 8 * p.nt[c.op] < mg_op_n(p, c.op)) return true;
        c.tile = static_cast<int>(blockIdx.x) - static_cast<int>(gridDim.x);
        if (++c.op == 6) {
            c.op = 0;
            ++c.l;
        }
    }
}
// warp 0: arm the slot's mbarrier and issue one bulk copy per weight row of the job
__device__ __forceinline__ void mg_fetch(const MegaParams& p, const MgCursor& c, unsigned char* slot, uint64_t* bar) {
    const int lane = threadIdx.x & 31;
    const int N = mg_op_n(p, c.op), K = mg_op_k(p, c.op), nt = p.nt[c.op];
    const int n0 = c.tile * 8 * nt;
    const int rows = min(8 * nt, N - n0);
    const int w_stride = K * 2 + kMgWPad;
    const __nv_bfloat16* W = p.layers[c.l].op[c.op].W;
    if (lane == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(rows) * K * 2);
    __syncwarp();
    for (int r = lane; r < rows; r += 32)
        mg_bulk_g2s(slot + static_cast<size_t>(r) * w_stride, W + static_cast<long long>(n0 + r) * K, static_cast<uint32_t>(K) * 2, bar);
}

struct MgSmem {
    unsigned char* slot[2];
    float* red;                              // [16 warps][NT][16][8]
    float* fresh;                            // EPI 2: [NT][MP][8] updated rows of this tile
    float* mean;                             // [64]
    float* rstd;                             // [64]
    uint64_t* bar;                           // [2]
    float* attn;                             // [kMgGroups][attention scratch]
};
constexpr int kMgAttnFloats = 64 + 4 + 4 + 4 * 64;      // s_q, s_p, s_red, s_out per group

// One linear-layer tile: out[M][8 NT] = epilogue(A[M][K] W[8 NT][K]^T).  IN_LN 0: bf16 activations; 2: folded LayerNorm
// (A = bf16 copy of the residual rows, statistics from `stats`).  EPI 0: fp32; 1: GELU -> bf16; 2: in-place residual update
// + bf16 copy + per-8-feature row-statistics partials.  Same arithmetic as gemv16_kernel (gemv.cu).
template <int IN_LN, int EPI, int NT, int MT>
__device__ __noinline__ void mg_linear_tile(const MegaParams& p, const MegaOp& op, int N, int K, int tile, const __nv_bfloat16* A,
                                               int stats_parts, float* out_f32, __nv_bfloat16* out_bf16, const MgSmem& sm,
                                               const unsigned char* w_s, uint64_t* wbar, uint32_t wparity) {
    constexpr int KS = kMgWarps / MT;
    constexpr int MP = 16 * MT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int mt = warp / KS, kw = warp % KS, rb = mt * 16;
    const int n0 = tile * 8 * NT;
    const int nkb = K >> 5;
    const int w_stride = K * 2 + kMgWPad;
    const int M = p.B;
    // diagnostics: CTA 0 / thread 0 accumulates the time between stages into trace[stage_base + 8 * IN_LN/EPI class + stage]
    unsigned long long* stage = (p.trace != nullptr && blockIdx.x == 0 && tid == 0) ? p.trace + p.stage_base + 8 * (IN_LN == 2 ? EPI : 2 + EPI) : nullptr;
    unsigned long long t_prev = stage ? mg_globaltimer() : 0ull;
    auto stamp = [&](int i) {
        if (stage) {
            const unsigned long long now = mg_globaltimer();
            stage[i] += now - t_prev;
            t_prev = now;
        }
    };
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.0f;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const int r_lo = rb + g, r_hi = rb + g + 8;
    const bool use_lo = r_lo < M && !(p.finished && p.finished[r_lo]);
    const bool use_hi = r_hi < M && !(p.finished && p.finished[r_hi]);
    const bool tile_live = __any_sync(0xffffffffu, use_lo || use_hi);
    const int my_blocks = (tile_live && kw < nkb) ? (nkb - kw + KS - 1) / KS : 0;
    const int n_batches = (my_blocks + kMgBatch - 1) / kMgBatch;
    uint4 a_lo[kMgBatch], a_hi[kMgBatch], n_lo[kMgBatch], n_hi[kMgBatch];
    auto load_batch = [&](int bi) {
#pragma unroll
        for (int i = 0; i < kMgBatch; ++i) {
            const int j = bi * kMgBatch + i;
            const int ko = (kw + j * KS) * 32 + t * 8;
            n_lo[i] = zero4;
            n_hi[i] = zero4;
            if (j < my_blocks) {
                if (use_lo) n_lo[i] = __ldcg(reinterpret_cast<const uint4*>(A + static_cast<long long>(r_lo) * K + ko));
                if (use_hi) n_hi[i] = __ldcg(reinterpret_cast<const uint4*>(A + static_cast<long long>(r_hi) * K + ko));
            }
        }
    };
    if (n_batches > 0) load_batch(0);
    if constexpr (IN_LN == 2) {
        constexpr int TPR = kMgThreads / MP;
        const int r = tid / TPR, sub = tid % TPR;
        float s1 = 0.0f, s2 = 0.0f;
        if (r < M) {
            for (int pp = sub; pp < stats_parts; pp += TPR) {
                const float2 v = __ldcg(reinterpret_cast<const float2*>(p.stats + (static_cast<long long>(pp) * MP + r) * 2));
                s1 += v.x;
                s2 += v.y;
            }
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (sub == 0) {
            const float mean = s1 / K;
            const float var = fmaxf(s2 / K - mean * mean, 0.0f);
            sm.mean[r] = mean;
            sm.rstd[r] = rsqrtf(var + 1e-5f);
            if (p.fold_flag != nullptr && blockIdx.x == 0 && r < M && !(p.finished && p.finished[r]) && mean * mean > 4.0f * var)
                *p.fold_flag = 1;
        }
        __syncthreads();
    }
    stamp(0);
    bool w_ready = false;
    for (int bi = 0; bi < n_batches; ++bi) {
#pragma unroll
        for (int i = 0; i < kMgBatch; ++i) {
            a_lo[i] = n_lo[i];
            a_hi[i] = n_hi[i];
        }
        if (bi + 1 < n_batches) load_batch(bi + 1);
        if (!w_ready) {
            stamp(1);
            mbar_wait(wbar, wparity);
            w_ready = true;
            stamp(2);
        }
#pragma unroll
        for (int i = 0; i < kMgBatch; ++i) {
            const int j = bi * kMgBatch + i;
            if (j < my_blocks) {
                const int kb = kw + j * KS;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint4 b = *reinterpret_cast<const uint4*>(w_s + static_cast<size_t>(nt * 8 + g) * w_stride + (kb * 32 + t * 8) * 2);
                    mg_mma(acc[nt], a_lo[i].x, a_hi[i].x, a_lo[i].y, a_hi[i].y, b.x, b.y);
                    mg_mma(acc[nt], a_lo[i].z, a_hi[i].z, a_lo[i].w, a_hi[i].w, b.z, b.w);
                }
            }
        }
    }
    if (!w_ready) mbar_wait(wbar, wparity);          // the slot is re-armed only after its copies have landed
    stamp(3);
    float* red_w = sm.red + warp * (NT * 128);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        red_w[nt * 128 + g * 8 + 2 * t] = acc[nt][0];
        red_w[nt * 128 + g * 8 + 2 * t + 1] = acc[nt][1];
        red_w[nt * 128 + (g + 8) * 8 + 2 * t] = acc[nt][2];
        red_w[nt * 128 + (g + 8) * 8 + 2 * t + 1] = acc[nt][3];
    }
    __syncthreads();                                  // every warp is done with the weight slot, too
    stamp(4);
    for (int idx = tid; idx < MT * NT * 128; idx += kMgThreads) {
        const int m = idx / (NT * 128), rem = idx - m * (NT * 128);
        const int nt = rem >> 7, r = (rem & 127) >> 3, c = rem & 7;
        const int row = m * 16 + r;
        const int n = n0 + nt * 8 + c;
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < KS; ++w) v += sm.red[(m * KS + w) * (NT * 128) + nt * 128 + r * 8 + c];
        const bool valid = row < M && n < N;
        if constexpr (IN_LN == 2) {
            if (valid) v = sm.rstd[row] * (v - sm.mean[row] * __ldg(op.c1 + n));
        }
        if (valid && op.bias) v += __ldg(op.bias + n);
        const bool store = valid && !(p.finished && p.finished[row]);
        const long long o = static_cast<long long>(row) * N + n;
        if constexpr (EPI == 0) {
            if (store) out_f32[o] = v;
        }
        if constexpr (EPI == 1) {
            if (store) out_bf16[o] = __float2bfloat16(gelu_fast(v));
        }
        if constexpr (EPI == 2) {
            float xn = 0.0f;
            if (valid) {
                xn = __ldcg(p.dx + o) + (store ? v : 0.0f);
                if (store) {
                    p.dx[o] = xn;
                    p.dxn[o] = __float2bfloat16(xn);
                }
            }
            sm.fresh[(nt * MP + row) * 8 + c] = xn;
        }
    }
    if constexpr (EPI == 2) {
        __syncthreads();
        // one (sum, sum of squares) partial per 8-feature n-tile and row: independent of how many n-tiles a CTA owns
        for (int idx = tid; idx < NT * MP; idx += kMgThreads) {
            const int nt = idx / MP, row = idx - nt * MP;
            if (n0 + nt * 8 < N) {
                float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float v = sm.fresh[(nt * MP + row) * 8 + c];
                    s1 += v;
                    s2 = fmaf(v, v, s2);
                }
                *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(n0 / 8 + nt) * MP + row) * 2) = make_float2(s1, s2);
            }
        }
    }
    stamp(5);
    __syncthreads();                                  // red / fresh / mean / rstd are reused by the CTA's next job
    stamp(6);
}

template <int IN_LN, int EPI, int MT>
__device__ __forceinline__ void mg_linear_nt(const MegaParams& p, const MegaOp& op, int nt, int N, int K, int tile, const __nv_bfloat16* A,
                                             int stats_parts, float* out_f32, __nv_bfloat16* out_bf16, const MgSmem& sm,
                                             const unsigned char* w_s, uint64_t* wbar, uint32_t wparity) {
    switch (nt) {
        case 1: mg_linear_tile<IN_LN, EPI, 1, MT>(p, op, N, K, tile, A, stats_parts, out_f32, out_bf16, sm, w_s, wbar, wparity); break;
        case 2: mg_linear_tile<IN_LN, EPI, 2, MT>(p, op, N, K, tile, A, stats_parts, out_f32, out_bf16, sm, w_s, wbar, wparity); break;
        case 3: mg_linear_tile<IN_LN, EPI, 3, MT>(p, op, N, K, tile, A, stats_parts, out_f32, out_bf16, sm, w_s, wbar, wparity); break;
        default: mg_linear_tile<IN_LN, EPI, 4, MT>(p, op, N, K, tile, A, stats_parts, out_f32, out_bf16, sm, w_s, wbar, wparity); break;
    }
}

// One (row, head) attention unit on a 128-thread group: decode_attention_kernel's body (decode.cu) with the
// projection inputs read from the fp32 `proj` rows.  self: append this position's k, v to the cache first.
__device__ __noinline__ void mg_attention_unit(const MegaParams& p, bool self_mode, int layer, int b, int h, int pos, int grp,
                                                  float* scratch) {
    const int tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;
    float* s_q = scratch;
    float* s_p = scratch + 64;
    float* s_red = scratch + 68;
    float* s_out = scratch + 72;                      // [4][64]
    const int d = p.d;
    const __nv_bfloat16 *K, *V;
    int n_keys;
    const int ld = self_mode ? 3 * d : d;
    const float* row = p.proj + static_cast<long long>(b) * ld;
    if (self_mode) {
        const long long blk = (static_cast<long long>(layer) * p.B + b) * p.H * p.tmax * 64 + static_cast<long long>(h) * p.tmax * 64;
        n_keys = pos + 1;
        if (tid < 64) {
            p.k_cache[blk + static_cast<long long>(pos) * 64 + tid] = __float2bfloat16(__ldcg(row + d + h * 64 + tid));
        } else {
            const int e = tid - 64;
            p.v_cache[blk + static_cast<long long>(pos) * 64 + e] = __float2bfloat16(__ldcg(row + 2 * d + h * 64 + e));
        }
        K = p.k_cache + blk;
        V = p.v_cache + blk;
    } else {
        const long long per_head = static_cast<long long>(p.T) * 64;
        const long long blk = static_cast<long long>(b / p.kv_div) * p.L * 2 * p.H * per_head + static_cast<long long>(h) * per_head;
        K = p.cross_kv + blk + (static_cast<long long>(layer) * 2 + 0) * p.H * per_head;
        V = p.cross_kv + blk + (static_cast<long long>(layer) * 2 + 1) * p.H * per_head;
        n_keys = p.T;
    }
    if (tid < 64) s_q[tid] = __bfloat162float(__float2bfloat16(__ldcg(row + h * 64 + tid)));
    mg_group_sync(grp);
    const int sub = lane >> 3, ch = lane & 7;
    float qv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = s_q[ch * 8 + i];
    float m_run = -INFINITY, l_run = 0.0f;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    constexpr int kStep = 4 * 4;
    for (int j0 = warp * 4; j0 < n_keys; j0 += 2 * kStep) {
        uint4 kraw[2], vraw[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = j0 + u * kStep + sub;
            ok[u] = j < n_keys;
            if (ok[u]) {
                const long long off = static_cast<long long>(j) * 64 + ch * 8;
                kraw[u] = *reinterpret_cast<const uint4*>(K + off);
                vraw[u] = *reinterpret_cast<const uint4*>(V + off);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float dot = 0.0f;
            if (ok[u]) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kraw[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __bfloat1622float2(h2[i]);
                    dot = fmaf(f.x, qv[2 * i], dot);
                    dot = fmaf(f.y, qv[2 * i + 1], dot);
                }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            if (ok[u]) {
                const float m_new = fmaxf(m_run, dot);
                const float scale = __expf(m_run - m_new);
                const float pj = __expf(dot - m_new);
                l_run = fmaf(l_run, scale, pj);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vraw[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __bfloat1622float2(h2[i]);
                    acc[2 * i] = fmaf(acc[2 * i], scale, pj * f.x);
                    acc[2 * i + 1] = fmaf(acc[2 * i + 1], scale, pj * f.y);
                }
                m_run = m_new;
            }
        }
    }
    float m_w = fmaxf(m_run, __shfl_xor_sync(0xffffffffu, m_run, 8));
    m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, 16));
    const float sc = (m_run == -INFINITY) ? 0.0f : __expf(m_run - m_w);
    l_run *= sc;
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 8);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 16);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] *= sc;
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (lane == 0) {
        s_red[warp] = m_w;
        s_p[warp] = l_run;
    }
    if (sub == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_out[warp * 64 + ch * 8 + i] = acc[i];
    }
    mg_group_sync(grp);
    if (tid < 64) {
        float gmax = s_red[0];
#pragma unroll
        for (int wi = 1; wi < 4; ++wi) gmax = fmaxf(gmax, s_red[wi]);
        float o = 0.0f, gsum = 0.0f;
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            const float f = (s_red[wi] == -INFINITY) ? 0.0f : __expf(s_red[wi] - gmax);
            o = fmaf(s_out[wi * 64 + tid], f, o);
            gsum = fmaf(s_p[wi], f, gsum);
        }
        p.datt[static_cast<long long>(b) * d + h * 64 + tid] = __float2bfloat16(o / gsum);
    }
    mg_group_sync(grp);                               // the scratch is reused by the group's next unit
}

template <int MT>
__global__ void __launch_bounds__(kMgThreads, 1) decode_mega_kernel(const MegaParams p) {
    constexpr int MP = 16 * MT;
    extern __shared__ __align__(128) unsigned char mg_smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ float s_mean[64], s_rstd[64];
    __shared__ float s_part[2][8];
    MgSmem sm;
    sm.slot[0] = mg_smem;
    sm.slot[1] = mg_smem + kMgSlotBytes;
    sm.red = reinterpret_cast<float*>(mg_smem + 2 * kMgSlotBytes);
    sm.fresh = sm.red + kMgWarps * kMgMaxNT * 128;
    sm.attn = sm.fresh + kMgMaxNT * 64 * 8;
    sm.mean = s_mean;
    sm.rstd = s_rstd;
    sm.bar = s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    // weight ring: the first two jobs' tiles are requested before the dependency wait (weights are constants)
    MgCursor fetch{0, 0, static_cast<int>(blockIdx.x) - static_cast<int>(gridDim.x)};
    bool more = mg_advance(p, fetch);
    int fetched = 0;
    if (warp == 0) {
        for (int s = 0; s < 2 && more; ++s) {
            mg_fetch(p, fetch, sm.slot[s], &s_bar[s]);
            ++fetched;
            more = mg_advance(p, fetch);
        }
    }
    pdl_launch_dependents();
    pdl_wait();
    const int B = p.B, d = p.d, H = p.H;
    const int pos = *p.step_ptr;
    if (p.trace != nullptr && blockIdx.x == 0 && tid == 0) p.trace[0] = mg_globaltimer();
    unsigned int barrier = 0;
    // ---- phase 0: embedding + exact row statistics (embed_kernel + row_stats_kernel, one row per CTA at a time)
    for (int r = blockIdx.x; r < B; r += gridDim.x) {
        const int tok = p.next_token[r];
        const __nv_bfloat16* e = p.emb + static_cast<size_t>(tok) * d;
        const float* pe = p.pos_emb + static_cast<size_t>(pos) * d;
        float s1 = 0.0f, s2 = 0.0f;
        if (tid < 256) {
            for (int i = tid; i < d; i += 256) {
                const float v = __bfloat162float(e[i]) + pe[i];
                p.dx[static_cast<size_t>(r) * d + i] = v;
                p.dxn[static_cast<size_t>(r) * d + i] = __float2bfloat16(v);
                s1 += v;
                s2 = fmaf(v, v, s2);
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if ((tid & 31) == 0) {
                s_part[0][tid >> 5] = s1;
                s_part[1][tid >> 5] = s2;
            }
        }
        __syncthreads();
        if (tid == 0) {
            float a = 0.0f, b2 = 0.0f;
            for (int w = 0; w < 8; ++w) {
                a += s_part[0][w];
                b2 += s_part[1][w];
            }
            p.stats[r * 2] = a;                        // [1 part][MP][2]
            p.stats[r * 2 + 1] = b2;
        }
        __syncthreads();
    }
    mg_grid_sync(p.sync, ++barrier, p.trace);

    MgCursor cur{0, 0, static_cast<int>(blockIdx.x) - static_cast<int>(gridDim.x)};
    bool have = mg_advance(p, cur);
    int done = 0;                                      // jobs computed so far: job i lives in slot i & 1, parity (i >> 1) & 1
    int stats_parts = 1;
    const int grp = tid >> 7;
    float* scratch = sm.attn + grp * kMgAttnFloats;
    for (int l = 0; l < p.L; ++l) {
        const MegaLayerDev& ly = p.layers[l];
        for (int op = 0; op < 6; ++op) {
            const int N = mg_op_n(p, op), K = mg_op_k(p, op);
            while (have && cur.l == l && cur.op == op) {
                const int s = done & 1;
                const uint32_t parity = (done >> 1) & 1;
                const MegaOp& o = ly.op[op];
                switch (op) {
                    case 0: mg_linear_nt<2, 0, MT>(p, o, p.nt[op], N, K, cur.tile, p.dxn, stats_parts, p.proj, nullptr, sm, sm.slot[s], &s_bar[s], parity); break;
                    case 2: mg_linear_nt<2, 0, MT>(p, o, p.nt[op], N, K, cur.tile, p.dxn, stats_parts, p.proj, nullptr, sm, sm.slot[s], &s_bar[s], parity); break;
                    case 4: mg_linear_nt<2, 1, MT>(p, o, p.nt[op], N, K, cur.tile, p.dxn, stats_parts, nullptr, p.dff, sm, sm.slot[s], &s_bar[s], parity); break;
                    case 5: mg_linear_nt<0, 2, MT>(p, o, p.nt[op], N, K, cur.tile, p.dff, 0, nullptr, nullptr, sm, sm.slot[s], &s_bar[s], parity); break;
                    default: mg_linear_nt<0, 2, MT>(p, o, p.nt[op], N, K, cur.tile, p.datt, 0, nullptr, nullptr, sm, sm.slot[s], &s_bar[s], parity); break;
                }
                ++done;
                have = mg_advance(p, cur);
                if (warp == 0 && more) {               // the slot just drained takes the job two ahead
                    mg_fetch(p, fetch, sm.slot[s], &s_bar[s]);
                    ++fetched;
                    more = mg_advance(p, fetch);
                }
            }
            if (op == 1 || op == 3 || op == 5) stats_parts = d / 8;
            mg_grid_sync(p.sync, ++barrier, p.trace);
            if (op == 0 || op == 2) {                  // attention phase on the projection just written
                const bool self_mode = op == 0;
                for (int u = blockIdx.x * kMgGroups + grp; u < B * H; u += gridDim.x * kMgGroups) {
                    const int b = u / H, h = u - b * H;
                    if (p.finished && p.finished[b]) continue;
                    mg_attention_unit(p, self_mode, l, b, h, pos, grp, scratch);
                }
                mg_grid_sync(p.sync, ++barrier, p.trace);
            }
        }
    }
    (void)fetched;
    (void)MP;
    // the last CTA out resets the counters for the next launch (nobody spins any more: every CTA passed the last barrier)
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(p.sync + 1, 1u) == gridDim.x - 1) {
            p.sync[0] = 0u;
            p.sync[1] = 0u;
            p.sync[32] = 0u;
            __threadfence();
        }
    }
}

static size_t mega_smem_bytes() {
    return 2 * static_cast<size_t>(kMgSlotBytes) + sizeof(float) * (kMgWarps * kMgMaxNT * 128 + kMgMaxNT * 64 * 8 + kMgGroups * kMgAttnFloats);
}

int mega_pick_nt(int N, int K, int n_ctas) {
    int nt = std::min(kMgMaxNT, std::max(1, ceil_div(ceil_div(N, 8), n_ctas)));
    while (nt > 1 && static_cast<size_t>(8 * nt) * (static_cast<size_t>(K) * 2 + kMgWPad) > static_cast<size_t>(kMgSlotBytes)) --nt;
    return nt;
}

bool mega_supported(int d, int F) {
    auto fits = [](int K) { return static_cast<size_t>(8) * (static_cast<size_t>(K) * 2 + kMgWPad) <= static_cast<size_t>(kMgSlotBytes); };
    return d % 64 == 0 && F % 32 == 0 && fits(d) && fits(F) && d / 8 <= 256;
}

int decode_layers_mega(const MegaArgs& a, cudaStream_t stream) {
    WSB_REQUIRE(a.B >= 1 && a.B <= 64, "the persistent decode kernel handles at most 64 rows");
    WSB_REQUIRE(mega_supported(a.d, a.F), "model width not supported by the persistent decode kernel");
    static PerDeviceOnce once;
    static int num_sms[64] = {0};
    int dev = 0;
    if (once.need(&dev)) {
        WSB_REQUIRE(dev < 64, "device index");
        const int smem = static_cast<int>(mega_smem_bytes());
        WSB_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        WSB_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        WSB_CHECK_CUDA(cudaFuncSetAttribute(decode_mega_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int sms = 0, occ = 0;
        WSB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        WSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decode_mega_kernel<4>, kMgThreads, smem));
        WSB_REQUIRE(occ >= 1, "the persistent decode kernel does not fit on an SM");
        num_sms[dev] = sms;
        once.mark(dev);
    }
    MegaParams p;
    p.layers = static_cast<const MegaLayerDev*>(a.layers_dev);
    p.L = a.L;
    p.d = a.d;
    p.F = a.F;
    p.H = a.H;
    p.T = a.T;
    p.tmax = a.tmax;
    p.B = a.B;
    p.next_token = a.next_token;
    p.step_ptr = a.step_ptr;
    p.emb = a.emb;
    p.pos_emb = a.pos_emb;
    p.dx = a.dx;
    p.dxn = a.dxn;
    p.stats = a.stats;
    p.proj = a.proj;
    p.datt = a.datt;
    p.dff = a.dff;
    p.k_cache = a.k_cache;
    p.v_cache = a.v_cache;
    p.cross_kv = a.cross_kv;
    p.finished = a.finished;
    p.kv_div = a.kv_div < 1 ? 1 : a.kv_div;
    p.sync = a.sync;
    p.fold_flag = a.fold_flag;
    p.trace = a.trace;
    p.stage_base = 2 * (2 + 10 * a.L) + 2;
    const int grid = num_sms[dev];
    const int Ns[6] = {3 * a.d, a.d, a.d, a.d, a.F, a.d};
    const int Ks[6] = {a.d, a.d, a.d, a.d, a.d, a.F};
    for (int i = 0; i < 6; ++i) p.nt[i] = mega_pick_nt(Ns[i], Ks[i], grid);
    const size_t smem = mega_smem_bytes();
    if (a.B <= 16) WSB_CHECK_CUDA(launch_kernel(decode_mega_kernel<1>, dim3(grid), dim3(kMgThreads), smem, stream, p));
    else if (a.B <= 32) WSB_CHECK_CUDA(launch_kernel(decode_mega_kernel<2>, dim3(grid), dim3(kMgThreads), smem, stream, p));
    else WSB_CHECK_CUDA(launch_kernel(decode_mega_kernel<4>, dim3(grid), dim3(kMgThreads), smem, stream, p));
    count_launch();
    return 0;
}

size_t mega_layer_table_bytes(int n_layers) { return sizeof(MegaLayerDev) * static_cast<size_t>(n_layers); }

void mega_fill_layer(void* host_table, int layer, const void* const W[6], const float* const bias[6], const float* const c1[6]) {
    MegaLayerDev* t = static_cast<MegaLayerDev*>(host_table) + layer;
    for (int i = 0; i < 6; ++i) {
        t->op[i].W = static_cast<const __nv_bfloat16*>(W[i]);
        t->op[i].bias = bias[i];
        t->op[i].c1 = c1[i];
    }
}

}  // namespace wsb
