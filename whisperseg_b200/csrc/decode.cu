// K5 -- kernels of the batched-over-windows greedy decode step that are not GEMMs:
//   * single-query attention over a head-major K/V block (self-attention over the growing cache with
//     fused cache append, and cross-attention over the 500 encoder frames)  -- HBM-bound streaming;
//   * the arg-max / logits-processor tail: combine the per-tile partial arg-max of the tied output
//     projection, apply finished-row masking, record the token, update the finished flags.
//
// Replaces, per generated token, HF WhisperDecoderLayer self/cross attention with its KV cache
// (modeling_whisper.py:417-506, 241-357) and the greedy branch of generate() with the
// SuppressTokens / SuppressTokensAtBegin processors (generation_whisper.py:1774-1812) -- the
// suppression masks are additive -inf vectors fused into the projection GEMM's epilogue.
#include "common.cuh"
#include "wsb_internal.h"
#include "decode.h"

#include <cstdlib>

namespace wsb {

constexpr int kDaThreads = 128;
constexpr int kDaWideThreads = 512;
constexpr int kDaMaxKeys = 512;

// 512 threads per (row, head) unit while all units fit on the device at once (4 such CTAs per SM); WSB_ATTN_THREADS=128
// pins the narrow variant (the two variants reduce in different orders: tests that compare batch sizes bit for bit pin it)
static bool decode_attention_wide(int B, int n_heads) {
    const char* e = std::getenv("WSB_ATTN_THREADS");
    const bool pinned = e != nullptr && std::atoi(e) == 128;
    return !pinned && static_cast<long long>(B) * n_heads <= 592;
}

// q: bf16 [B][q_ld] (+ head*64), already scaled.  K/V blocks: bf16 [n_keys][64] contiguous per (b, head).
// mode 0 (cross): K at kv + ((b*kv_heads_total + k_slot)*T)*64, n_keys = T, fixed.
// mode 1 (self) : cache [B][H][Tmax][64] for K and V; the new k,v (from the packed qkv row) are
//                 appended at position *step_ptr + pos_offset before attending over [0, pos].
struct DaParams {
    const __nv_bfloat16* q;
    long long q_ld;
    const __nv_bfloat16* k_base;
    const __nv_bfloat16* v_base;
    __nv_bfloat16* k_cache;       // self mode only (same memory as k_base)
    __nv_bfloat16* v_cache;
    const __nv_bfloat16* new_k;   // self mode: pointer to k part of the packed row (b stride q_ld)
    const __nv_bfloat16* new_v;
    long long bh_stride;          // elements between consecutive (b,h) blocks
    long long b_stride;           // elements between consecutive b
    int n_keys_fixed;             // cross: T
    const int* step_ptr;          // self: current step (device)
    int pos_offset;               // self: position = pos_offset + *step_ptr
    const unsigned char* finished;
    __nv_bfloat16* out;           // [B][d]
    int d;
    int self_mode;
    // optional fused split-K second phase: q (and the new k, v in self mode) = bias + sum of fp32 partial planes
    const float* part;            // [splits][B][part_ld] or null (then q / new_k / new_v bf16 are read)
    int splits;
    long long split_stride;
    int part_ld;
    const float* bias;            // [part_ld]
    // beam search: rows sharing a cross K/V block; self-attention ancestry tables [2][B][anc_ld]
    int kv_div;
    const int* anc;
    int anc_ld;
    long long anc_buf_stride;
};

// kThreads = 128, or 512 when there are so few (row, head) units that a 128-thread CTA per unit leaves the HBM pipe
// empty: a unit streams its 128 KB of K/V with (threads x 64 B) in flight, i.e. ~8 GB/s per 128-thread CTA at ~1 us
// of latency -- 16-20 us per launch at <= 16 rows whatever the row count (r2 mega trace: cross-attention was 37 us of a
// 110 us layer at 8 rows).  Four times the threads per unit = four times the bytes in flight.
template <bool kAnc, int kThreads>
__global__ void __launch_bounds__(kThreads) decode_attention_kernel(const DaParams p) {
    constexpr int kDaThreads = kThreads;
    const int h = blockIdx.x, b = blockIdx.y;
    pdl_wait();
    pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ float s_q[64];
    __shared__ int s_src[kAnc ? kDaMaxKeys : 1];
    __shared__ float s_p[kDaThreads / 32];
    __shared__ float s_red[kDaThreads / 32];
    __shared__ float s_out[kDaThreads / 32][64];

    const long long blk = static_cast<long long>(p.self_mode ? b : b / p.kv_div) * p.b_stride + static_cast<long long>(h) * p.bh_stride;
    const __nv_bfloat16* K = p.k_base + blk;
    const __nv_bfloat16* V = p.v_base + blk;
    int n_keys;
    // value of column `col` of this row's projection output: bf16 activation, or bias + split-K planes.  The planes are
    // fetched together (one L2 round trip, not one per plane) and added in plane order.
    auto proj = [&](int col, const __nv_bfloat16* act) -> float {
        if (p.part == nullptr) return __bfloat162float(act[static_cast<long long>(b) * p.q_ld + h * 64 + (col & 63)]);
        const float* src = p.part + static_cast<long long>(b) * p.part_ld + col;
        float t[16];
#pragma unroll
        for (int sp = 0; sp < 16; ++sp) t[sp] = sp < p.splits ? src[sp * p.split_stride] : 0.0f;
        float acc = p.bias ? __ldg(p.bias + col) : 0.0f;
#pragma unroll
        for (int sp = 0; sp < 16; ++sp)
            if (sp < p.splits) acc += t[sp];
        return acc;
    };
    // the projection inputs are requested together with the finished flag; a finished row leaves before it writes anything
    const bool fin = p.finished && p.finished[b];
    float new_kv = 0.0f, new_q = 0.0f;
    int pos = 0;
    if (p.self_mode) {
        pos = p.pos_offset + *p.step_ptr;
        if (tid < 64) new_kv = proj(p.d + h * 64 + tid, p.new_k);
        else if (tid < 128) new_kv = proj(2 * p.d + h * 64 + (tid - 64), p.new_v);
    }
    if (tid < 64) new_q = proj(h * 64 + tid, p.q);
    if (fin) return;
    if (p.self_mode) {
        n_keys = pos + 1;
        if constexpr (kAnc) {
            const int* a = p.anc + (pos & 1) * p.anc_buf_stride + static_cast<long long>(b) * p.anc_ld;
            for (int j = tid; j <= pos; j += kDaThreads) s_src[j] = (j == pos) ? b : a[j];
        }
        if (tid < 64) p.k_cache[blk + static_cast<long long>(pos) * 64 + tid] = __float2bfloat16(new_kv);
        else if (tid < 128) p.v_cache[blk + static_cast<long long>(pos) * 64 + (tid - 64)] = __float2bfloat16(new_kv);
    } else {
        n_keys = p.n_keys_fixed;
    }
    // q is rounded to bf16 like the stand-alone second phase would, so both paths give identical results
    if (tid < 64) s_q[tid] = __bfloat162float(__float2bfloat16(new_q));
    __syncthreads();                                   // also orders the cache append before the reads below

    // Single pass over the keys with an online softmax: a warp covers 4 keys per step (lane = (key % 4) * 8 +
    // 16-byte chunk, so every load instruction reads 512 contiguous bytes), K and V rows of two steps are in
    // flight together, and each 8-lane group keeps its own running (max, sum, 8-dim accumulator).
    const int sub = lane >> 3, ch = lane & 7;
    float qv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = s_q[ch * 8 + i];
    float m_run = -INFINITY, l_run = 0.0f;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    constexpr int kStep = (kDaThreads / 32) * 4;           // keys per CTA step
    for (int j0 = warp * 4; j0 < n_keys; j0 += 2 * kStep) {
        uint4 kraw[2], vraw[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = j0 + u * kStep + sub;
            ok[u] = j < n_keys;
            if (ok[u]) {
                long long off = static_cast<long long>(j) * 64 + ch * 8;
                if constexpr (kAnc) off += static_cast<long long>(s_src[j] - b) * p.b_stride;
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
                const float scale = __expf(m_run - m_new);      // exp(-inf) = 0 on the first key
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
    // merge the 4 key sub-groups of the warp, then the warps
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
        for (int i = 0; i < 8; ++i) s_out[warp][ch * 8 + i] = acc[i];
    }
    __syncthreads();
    if (tid < 64) {
        float gmax = s_red[0];
#pragma unroll
        for (int wi = 1; wi < kDaThreads / 32; ++wi) gmax = fmaxf(gmax, s_red[wi]);
        float o = 0.0f, gsum = 0.0f;
#pragma unroll
        for (int wi = 0; wi < kDaThreads / 32; ++wi) {
            const float f = (s_red[wi] == -INFINITY) ? 0.0f : __expf(s_red[wi] - gmax);
            o = fmaf(s_out[wi][tid], f, o);
            gsum = fmaf(s_p[wi], f, gsum);
        }
        p.out[static_cast<long long>(b) * p.d + h * 64 + tid] = __float2bfloat16(o / gsum);
    }
}

int decode_self_attention(const __nv_bfloat16* qkv, const SplitkInput* part, int d, __nv_bfloat16* k_cache,
                          __nv_bfloat16* v_cache, int t_max, const int* step_ptr, int pos_offset,
                          const unsigned char* finished, __nv_bfloat16* out, int B, int n_heads, cudaStream_t stream,
                          const int* anc, int anc_ld) {
    WSB_REQUIRE(t_max <= kDaMaxKeys, "self-attention cache longer than 512 positions");
    if (B <= 0) return 0;
    DaParams p;
    p.q = qkv;
    p.q_ld = 3LL * d;
    p.k_base = k_cache;
    p.v_base = v_cache;
    p.k_cache = k_cache;
    p.v_cache = v_cache;
    p.new_k = qkv + d;
    p.new_v = qkv + 2 * d;
    p.bh_stride = static_cast<long long>(t_max) * 64;
    p.b_stride = static_cast<long long>(n_heads) * t_max * 64;
    p.n_keys_fixed = 0;
    p.step_ptr = step_ptr;
    p.pos_offset = pos_offset;
    p.finished = finished;
    p.out = out;
    p.d = d;
    p.self_mode = 1;
    p.part = part ? part->planes : nullptr;
    p.splits = part ? part->splits : 0;
    p.split_stride = part ? part->split_stride : 0;
    p.part_ld = 3 * d;
    p.bias = part ? part->bias : nullptr;
    p.kv_div = 1;
    p.anc = anc;
    p.anc_ld = anc_ld;
    p.anc_buf_stride = static_cast<long long>(B) * anc_ld;
    if (anc) {
        WSB_CHECK_CUDA(launch_kernel(decode_attention_kernel<true, kDaThreads>, dim3(n_heads, B), dim3(kDaThreads), 0, stream, p));
        count_launch();
        return 0;
    }
    if (decode_attention_wide(B, n_heads))
        WSB_CHECK_CUDA(launch_kernel(decode_attention_kernel<false, kDaWideThreads>, dim3(n_heads, B), dim3(kDaWideThreads), 0, stream, p));
    else
        WSB_CHECK_CUDA(launch_kernel(decode_attention_kernel<false, kDaThreads>, dim3(n_heads, B), dim3(kDaThreads), 0, stream, p));
    count_launch();
    return 0;
}

int decode_cross_attention(const __nv_bfloat16* q, const SplitkInput* part, int d, const __nv_bfloat16* cross_kv, int layer,
                           int n_layers, int T, const unsigned char* finished, __nv_bfloat16* out, int B, int n_heads,
                           cudaStream_t stream, int kv_div) {
    WSB_REQUIRE(T <= kDaMaxKeys, "cross-attention over more than 512 encoder positions");
    if (B <= 0) return 0;
    // cross_kv layout (written by the head-major GEMM epilogue): [b][layer][k|v][head][T][64]
    DaParams p;
    p.q = q;
    p.q_ld = d;
    const long long per_head = static_cast<long long>(T) * 64;
    p.k_base = cross_kv + (static_cast<long long>(layer) * 2 + 0) * n_heads * per_head;
    p.v_base = cross_kv + (static_cast<long long>(layer) * 2 + 1) * n_heads * per_head;
    p.k_cache = nullptr;
    p.v_cache = nullptr;
    p.new_k = nullptr;
    p.new_v = nullptr;
    p.bh_stride = per_head;
    p.b_stride = static_cast<long long>(n_layers) * 2 * n_heads * per_head;
    p.n_keys_fixed = T;
    p.step_ptr = nullptr;
    p.pos_offset = 0;
    p.finished = finished;
    p.out = out;
    p.d = d;
    p.self_mode = 0;
    p.part = part ? part->planes : nullptr;
    p.splits = part ? part->splits : 0;
    p.split_stride = part ? part->split_stride : 0;
    p.part_ld = d;
    p.bias = part ? part->bias : nullptr;
    p.kv_div = kv_div < 1 ? 1 : kv_div;
    p.anc = nullptr;
    p.anc_ld = 0;
    p.anc_buf_stride = 0;
    if (decode_attention_wide(B, n_heads))
        WSB_CHECK_CUDA(launch_kernel(decode_attention_kernel<false, kDaWideThreads>, dim3(n_heads, B), dim3(kDaWideThreads), 0, stream, p));
    else
        WSB_CHECK_CUDA(launch_kernel(decode_attention_kernel<false, kDaThreads>, dim3(n_heads, B), dim3(kDaThreads), 0, stream, p));
    count_launch();
    return 0;
}

// ---------------------------------------------------------------------------- prompt prefill
// The prompt ([<|startoftranscript|>, <|en|>, <|notimestamps|>], reference model.py:655-661) is the same for every row, and the
// reference's generate() runs it through the decoder in one forward pass.  Here too: the P prompt positions of all B rows
// form P * B virtual rows (row p * B + b) for the linear layers, and the two attention kernels below serve the P positions
// of a row together -- the cross-attention K/V block of a row (128 KB per head) is streamed once instead of P times.
__global__ void embed_prefill_kernel(const int* __restrict__ prompt, const int* __restrict__ forced, int forced_ld,
                                     const __nv_bfloat16* __restrict__ emb, const float* __restrict__ pos_emb,
                                     float* __restrict__ x, int B, int d) {
    const int r = blockIdx.x, p = r / B, b = r - p * B;
    const int tok = forced ? forced[static_cast<long long>(b) * forced_ld + p] : prompt[p];
    const __nv_bfloat16* e = emb + static_cast<size_t>(tok) * d;
    const float* pe = pos_emb + static_cast<size_t>(p) * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[static_cast<size_t>(r) * d + i] = __bfloat162float(e[i]) + pe[i];
}

int embed_prefill(const int* prompt_dev, const int* forced, int forced_ld, const __nv_bfloat16* emb, const float* pos_emb,
                  float* x, int B, int P, int d, cudaStream_t stream) {
    if (B <= 0 || P <= 0) return 0;
    WSB_CHECK_CUDA(launch_kernel(embed_prefill_kernel, dim3(B * P), dim3(256), 0, stream, prompt_dev, forced, forced_ld, emb, pos_emb, x, B, d));
    count_launch();
    return 0;
}

constexpr int kPrefillMaxP = 4;

__device__ __forceinline__ float prefill_proj(const DaParams& p, long long row, int col) {
    const float* src = p.part + row * p.part_ld + col;
    float t[16];
#pragma unroll
    for (int sp = 0; sp < 16; ++sp) t[sp] = sp < p.splits ? src[sp * p.split_stride] : 0.0f;
    float acc = p.bias ? __ldg(p.bias + col) : 0.0f;
#pragma unroll
    for (int sp = 0; sp < 16; ++sp)
        if (sp < p.splits) acc += t[sp];
    return acc;
}

// grid (H, B), 128 threads.  P <= 4 positions: q, k, v rounded to bf16 exactly like the per-position kernel (k, v through
// the cache, q in registers), scores and softmax in fp32.
__global__ void __launch_bounds__(128) prefill_self_attention_kernel(const DaParams p, int B, int P) {
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float s_q[kPrefillMaxP][64], s_k[kPrefillMaxP][64], s_v[kPrefillMaxP][64];
    __shared__ float s_w[kPrefillMaxP][kPrefillMaxP];
    const long long blk = static_cast<long long>(b) * p.b_stride + static_cast<long long>(h) * p.bh_stride;
    for (int idx = tid; idx < P * 64; idx += 128) {
        const int pos = idx >> 6, e = idx & 63;
        const long long row = static_cast<long long>(pos) * B + b;
        const __nv_bfloat16 kb = __float2bfloat16(prefill_proj(p, row, p.d + h * 64 + e));
        const __nv_bfloat16 vb = __float2bfloat16(prefill_proj(p, row, 2 * p.d + h * 64 + e));
        p.k_cache[blk + static_cast<long long>(pos) * 64 + e] = kb;
        p.v_cache[blk + static_cast<long long>(pos) * 64 + e] = vb;
        s_k[pos][e] = __bfloat162float(kb);
        s_v[pos][e] = __bfloat162float(vb);
        s_q[pos][e] = __bfloat162float(__float2bfloat16(prefill_proj(p, row, h * 64 + e)));
    }
    __syncthreads();
    if (tid < P * P) {                                   // score of (query position qp, key position kp <= qp)
        const int qp = tid / P, kp = tid - qp * P;
        float dot = -INFINITY;
        if (kp <= qp) {
            dot = 0.0f;
            for (int e = 0; e < 64; ++e) dot = fmaf(s_k[kp][e], s_q[qp][e], dot);
        }
        s_w[qp][kp] = dot;
    }
    __syncthreads();
    if (tid < P) {
        float mx = -INFINITY;
        for (int kp = 0; kp <= tid; ++kp) mx = fmaxf(mx, s_w[tid][kp]);
        float sum = 0.0f;
        for (int kp = 0; kp < P; ++kp) {
            const float w = kp <= tid ? __expf(s_w[tid][kp] - mx) : 0.0f;
            s_w[tid][kp] = w;
            sum += w;
        }
        for (int kp = 0; kp < P; ++kp) s_w[tid][kp] /= sum;
    }
    __syncthreads();
    for (int idx = tid; idx < P * 64; idx += 128) {
        const int qp = idx >> 6, e = idx & 63;
        float o = 0.0f;
        for (int kp = 0; kp <= qp; ++kp) o = fmaf(s_w[qp][kp], s_v[kp][e], o);
        p.out[(static_cast<long long>(qp) * B + b) * p.d + h * 64 + e] = __float2bfloat16(o);
    }
}

int prefill_self_attention(const SplitkInput* part, int d, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int t_max,
                           __nv_bfloat16* out, int B, int P, int n_heads, cudaStream_t stream) {
    WSB_REQUIRE(part != nullptr && P >= 1 && P <= kPrefillMaxP && P <= t_max, "prefill: 1..4 prompt positions, projection planes");
    if (B <= 0) return 0;
    DaParams p = {};
    p.k_cache = k_cache;
    p.v_cache = v_cache;
    p.bh_stride = static_cast<long long>(t_max) * 64;
    p.b_stride = static_cast<long long>(n_heads) * t_max * 64;
    p.out = out;
    p.d = d;
    p.part = part->planes;
    p.splits = part->splits;
    p.split_stride = part->split_stride;
    p.part_ld = 3 * d;
    p.bias = part->bias;
    WSB_CHECK_CUDA(launch_kernel(prefill_self_attention_kernel, dim3(n_heads, B), dim3(128), 0, stream, p, B, P));
    count_launch();
    return 0;
}

// grid (H, B), 128 threads: NQ online-softmax states per thread, one pass over the row's K/V block; per query the
// arithmetic and its order are decode_attention_kernel's (128-thread form)
template <int NQ>
__global__ void __launch_bounds__(kDaThreads) prefill_cross_attention_kernel(const DaParams p, int B) {
    const int h = blockIdx.x, w = blockIdx.y;
    pdl_wait();
    pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ float s_q[NQ][64];
    __shared__ float s_p[NQ][kDaThreads / 32];
    __shared__ float s_red[NQ][kDaThreads / 32];
    __shared__ float s_out[NQ][kDaThreads / 32][64];
    const long long blk = static_cast<long long>(w) * p.b_stride + static_cast<long long>(h) * p.bh_stride;
    const __nv_bfloat16* K = p.k_base + blk;
    const __nv_bfloat16* V = p.v_base + blk;
    const int n_keys = p.n_keys_fixed;
    if (tid < 64) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi)
            s_q[qi][tid] = __bfloat162float(__float2bfloat16(prefill_proj(p, static_cast<long long>(qi) * B + w, h * 64 + tid)));
    }
    __syncthreads();
    const int sub = lane >> 3, ch = lane & 7;
    float qv[NQ][8], m_run[NQ], l_run[NQ], acc[NQ][8];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        m_run[qi] = -INFINITY;
        l_run[qi] = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            qv[qi][i] = s_q[qi][ch * 8 + i];
            acc[qi][i] = 0.0f;
        }
    }
    constexpr int kStep = (kDaThreads / 32) * 4;
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
            float kf[8], vf[8];
            if (ok[u]) {
                const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kraw[u]);
                const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vraw[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 a = __bfloat1622float2(k2[i]), c = __bfloat1622float2(v2[i]);
                    kf[2 * i] = a.x; kf[2 * i + 1] = a.y;
                    vf[2 * i] = c.x; vf[2 * i + 1] = c.y;
                }
            }
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
                float dot = 0.0f;
                if (ok[u]) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) dot = fmaf(kf[i], qv[qi][i], dot);
                }
                dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                if (ok[u]) {
                    const float m_new = fmaxf(m_run[qi], dot);
                    const float scale = __expf(m_run[qi] - m_new);
                    const float pj = __expf(dot - m_new);
                    l_run[qi] = fmaf(l_run[qi], scale, pj);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[qi][i] = fmaf(acc[qi][i], scale, pj * vf[i]);
                    m_run[qi] = m_new;
                }
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        float m_w = fmaxf(m_run[qi], __shfl_xor_sync(0xffffffffu, m_run[qi], 8));
        m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, 16));
        const float sc = (m_run[qi] == -INFINITY) ? 0.0f : __expf(m_run[qi] - m_w);
        float l = l_run[qi] * sc;
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a = acc[qi][i] * sc;
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            acc[qi][i] = a;
        }
        if (lane == 0) {
            s_red[qi][warp] = m_w;
            s_p[qi][warp] = l;
        }
        if (sub == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s_out[qi][warp][ch * 8 + i] = acc[qi][i];
        }
    }
    __syncthreads();
    if (tid < 64) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            float gmax = s_red[qi][0];
#pragma unroll
            for (int wi = 1; wi < kDaThreads / 32; ++wi) gmax = fmaxf(gmax, s_red[qi][wi]);
            float o = 0.0f, gsum = 0.0f;
#pragma unroll
            for (int wi = 0; wi < kDaThreads / 32; ++wi) {
                const float f = (s_red[qi][wi] == -INFINITY) ? 0.0f : __expf(s_red[qi][wi] - gmax);
                o = fmaf(s_out[qi][wi][tid], f, o);
                gsum = fmaf(s_p[qi][wi], f, gsum);
            }
            p.out[(static_cast<long long>(qi) * B + w) * p.d + h * 64 + tid] = __float2bfloat16(o / gsum);
        }
    }
}

int prefill_cross_attention(const SplitkInput* part, int d, const __nv_bfloat16* cross_kv, int layer, int n_layers, int T,
                            __nv_bfloat16* out, int B, int P, int n_heads, cudaStream_t stream) {
    WSB_REQUIRE(part != nullptr && P >= 1 && P <= kPrefillMaxP && T <= kDaMaxKeys, "prefill: 1..4 prompt positions, projection planes");
    if (B <= 0) return 0;
    DaParams p = {};
    const long long per_head = static_cast<long long>(T) * 64;
    p.k_base = cross_kv + (static_cast<long long>(layer) * 2 + 0) * n_heads * per_head;
    p.v_base = cross_kv + (static_cast<long long>(layer) * 2 + 1) * n_heads * per_head;
    p.bh_stride = per_head;
    p.b_stride = static_cast<long long>(n_layers) * 2 * n_heads * per_head;
    p.n_keys_fixed = T;
    p.out = out;
    p.d = d;
    p.part = part->planes;
    p.splits = part->splits;
    p.split_stride = part->split_stride;
    p.part_ld = d;
    p.bias = part->bias;
    const dim3 grid(n_heads, B);
    switch (P) {
        case 1: WSB_CHECK_CUDA(launch_kernel(prefill_cross_attention_kernel<1>, grid, dim3(kDaThreads), 0, stream, p, B)); break;
        case 2: WSB_CHECK_CUDA(launch_kernel(prefill_cross_attention_kernel<2>, grid, dim3(kDaThreads), 0, stream, p, B)); break;
        case 3: WSB_CHECK_CUDA(launch_kernel(prefill_cross_attention_kernel<3>, grid, dim3(kDaThreads), 0, stream, p, B)); break;
        default: WSB_CHECK_CUDA(launch_kernel(prefill_cross_attention_kernel<4>, grid, dim3(kDaThreads), 0, stream, p, B)); break;
    }
    count_launch();
    return 0;
}

// ---------------------------------------------------------------------------- arg-max tail
// one warp per row: reduce the per-tile partials (lowest index wins ties, like torch.argmax), then
// finished rows emit pad, the token is recorded and fed back, EOS marks the row finished.
// `*step_ptr` is the decoder position of the token just consumed; the token produced here lands in
// tokens_out[row][pos - out_offset].  With `forced` (teacher forcing, parity tests) the next input is
// forced[row][pos + 1] instead of the arg-max and rows never finish.
__global__ void argmax_finalize_kernel(const float* __restrict__ val, const int* __restrict__ idx, int n_tiles,
                                       int* __restrict__ tokens_out, int max_new, int out_offset,
                                       int* __restrict__ next_token, const int* __restrict__ forced, int forced_ld,
                                       unsigned char* __restrict__ finished, const int* __restrict__ step_ptr,
                                       int* __restrict__ n_active, int eos_id, int pad_id, int B,
                                       const int* __restrict__ row_map) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    const int pos = *step_ptr;
    if (warp < B) {
        float best = -INFINITY;
        int best_i = 0x7fffffff;
        for (int t = lane; t < n_tiles; t += 32) {
            const float v = val[static_cast<long long>(warp) * n_tiles + t];
            const int i = idx[static_cast<long long>(warp) * n_tiles + t];
            if (v > best || (v == best && i < best_i)) {
                best = v;
                best_i = i;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) {
                best = ov;
                best_i = oi;
            }
        }
        if (lane == 0) {
            const int slot = pos - out_offset;
            if (forced) {
                if (slot >= 0 && slot < max_new) tokens_out[static_cast<long long>(warp) * max_new + slot] = best_i;
                next_token[warp] = (pos + 1 < forced_ld) ? forced[static_cast<long long>(warp) * forced_ld + pos + 1] : pad_id;
            } else {
                const bool was_finished = finished[warp] != 0;
                const int tok = was_finished ? pad_id : best_i;
                const int orow = row_map ? row_map[warp] : warp;       // slot -> original window after compaction
                if (slot >= 0 && slot < max_new && !was_finished)
                    tokens_out[static_cast<long long>(orow) * max_new + slot] = tok;
                next_token[warp] = tok;
                if (!was_finished && tok == eos_id) {
                    finished[warp] = 1;
                    atomicSub(n_active, 1);
                }
            }
        }
    }
}
__global__ void step_increment_kernel(int* step_ptr) {
    pdl_wait();
    pdl_launch_dependents();
    *step_ptr += 1;
}

int step_increment(int* step_ptr, cudaStream_t stream) {
    WSB_CHECK_CUDA(launch_kernel(step_increment_kernel, dim3(1), dim3(1), 0, stream, step_ptr));
    count_launch();
    return 0;
}

int argmax_finalize(const float* val, const int* idx, int n_tiles, int* tokens_out, int max_new, int out_offset,
                    int* next_token, const int* forced, int forced_ld, unsigned char* finished, int* step_ptr,
                    int* n_active, int eos_id, int pad_id, int B, const int* row_map, cudaStream_t stream) {
    if (B <= 0) return 0;
    WSB_CHECK_CUDA(launch_kernel(argmax_finalize_kernel, dim3(ceil_div(B * 32, 256)), dim3(256), 0, stream, val, idx, n_tiles,
                                 tokens_out, max_new, out_offset, next_token, forced, forced_ld, finished,
                                 static_cast<const int*>(step_ptr), n_active, eos_id, pad_id, B, row_map));
    WSB_CHECK_CUDA(launch_kernel(step_increment_kernel, dim3(1), dim3(1), 0, stream, step_ptr));
    count_launch(2);
    return 0;
}

// prompt positions before the last one: no logits, just feed the next prompt token
__global__ void prefill_advance_kernel(int* __restrict__ next_token, const int* __restrict__ forced, int forced_ld,
                                       const int* __restrict__ prompt, int* __restrict__ step_ptr, int B) {
    pdl_wait();
    pdl_launch_dependents();
    const int pos = *step_ptr;
    for (int b = threadIdx.x; b < B; b += blockDim.x)
        next_token[b] = forced ? forced[static_cast<long long>(b) * forced_ld + pos + 1] : prompt[pos + 1];
    __syncthreads();
    if (threadIdx.x == 0) *step_ptr = pos + 1;
}

int prefill_advance(int* next_token, const int* forced, int forced_ld, const int* prompt_dev, int* step_ptr, int B,
                    cudaStream_t stream) {
    WSB_CHECK_CUDA(launch_kernel(prefill_advance_kernel, dim3(1), dim3(256), 0, stream, next_token, forced, forced_ld, prompt_dev,
                                 step_ptr, B));
    count_launch();
    return 0;
}

// ---------------------------------------------------------------------------- batch compaction
// When most windows have emitted EOS the per-position cost is dominated by work that scales with the
// batch dimension of the launch (split-K planes, reduce CTAs, GEMM m-tiles).  The still-active rows are
// then gathered into a small dense batch: their K/V caches and cross-attention K/V blocks are copied
// once (tens of MB per row) and the remaining hundreds of positions run on the compact batch.
__global__ void compact_plan_kernel(const unsigned char* __restrict__ fin_src, int b_src, const int* __restrict__ map_src,
                                    const int* __restrict__ tok_src, int b_dst, int* __restrict__ active_idx,
                                    int* __restrict__ map_dst, int* __restrict__ tok_dst,
                                    unsigned char* __restrict__ fin_dst) {
    if (threadIdx.x != 0) return;
    int n = 0;
    for (int r = 0; r < b_src && n < b_dst; ++r) {
        if (!fin_src[r]) {
            active_idx[n] = r;
            map_dst[n] = map_src ? map_src[r] : r;
            tok_dst[n] = tok_src[r];
            fin_dst[n] = 0;
            ++n;
        }
    }
    for (int i = n; i < b_dst; ++i) {
        active_idx[i] = -1;
        map_dst[i] = 0;
        tok_dst[i] = 0;
        fin_dst[i] = 1;                                  // padding slots are skipped by every kernel
    }
}

// grid (H, b_dst, L): copy positions [0, *step_ptr) of one (layer, slot, head) K and V block
__global__ void gather_self_cache_kernel(const __nv_bfloat16* __restrict__ k_src, const __nv_bfloat16* __restrict__ v_src,
                                         __nv_bfloat16* __restrict__ k_dst, __nv_bfloat16* __restrict__ v_dst,
                                         const int* __restrict__ active_idx, const int* __restrict__ step_ptr, int b_src,
                                         int b_dst, int n_heads, int t_max) {
    const int h = blockIdx.x, slot = blockIdx.y, l = blockIdx.z;
    const int r = active_idx[slot];
    if (r < 0) return;
    const int n_vec = *step_ptr * 8;                     // uint4 = 8 bf16; 64 dims = 8 vectors per position
    const long long so = ((static_cast<long long>(l) * b_src + r) * n_heads + h) * t_max * 64;
    const long long dofs = ((static_cast<long long>(l) * b_dst + slot) * n_heads + h) * t_max * 64;
    const uint4* ks = reinterpret_cast<const uint4*>(k_src + so);
    const uint4* vs = reinterpret_cast<const uint4*>(v_src + so);
    uint4* kd = reinterpret_cast<uint4*>(k_dst + dofs);
    uint4* vd = reinterpret_cast<uint4*>(v_dst + dofs);
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
        kd[i] = ks[i];
        vd[i] = vs[i];
    }
}

// grid (chunks, b_dst): copy one window's contiguous cross-attention K/V block [L][2][H][T][64]
__global__ void gather_cross_kv_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                       const int* __restrict__ active_idx, long long row_elems) {
    const int slot = blockIdx.y;
    const int r = active_idx[slot];
    if (r < 0) return;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * row_elems);
    uint4* d4 = reinterpret_cast<uint4*>(dst + static_cast<long long>(slot) * row_elems);
    const long long n_vec = row_elems / 8;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        d4[i] = s4[i];
}

int compact_decode_state(const CompactArgs& a, cudaStream_t stream) {
    compact_plan_kernel<<<1, 32, 0, stream>>>(a.fin_src, a.b_src, a.map_src, a.tok_src, a.b_dst, a.active_idx, a.map_dst,
                                              a.tok_dst, a.fin_dst);
    WSB_CHECK_CUDA(cudaGetLastError());
    gather_self_cache_kernel<<<dim3(a.n_heads, a.b_dst, a.n_layers), 256, 0, stream>>>(
        a.k_src, a.v_src, a.k_dst, a.v_dst, a.active_idx, a.step_ptr, a.b_src, a.b_dst, a.n_heads, a.t_max);
    WSB_CHECK_CUDA(cudaGetLastError());
    gather_cross_kv_kernel<<<dim3(256, a.b_dst), 256, 0, stream>>>(a.cross_src, a.cross_dst, a.active_idx, a.cross_row_elems);
    WSB_CHECK_CUDA(cudaGetLastError());
    count_launch(3);
    return 0;
}

}  // namespace wsb
