// Bandwidth-bound helper kernels: LayerNorm (fp32 residual stream -> bf16 GEMM operand), the conv1
// stem (80 -> d, k=3, pad=1, GELU) and decoder token/position embedding.
//
//   LayerNorm  : HF WhisperEncoderLayer / WhisperDecoderLayer nn.LayerNorm(eps=1e-5)
//                (modeling_whisper.py:361-414, 417-506)
//   conv1+GELU : HF WhisperEncoder.forward, modeling_whisper.py:619-620
//   embedding  : HF WhisperDecoder.forward, modeling_whisper.py:742-760 (embed_tokens + embed_positions)
#include "common.cuh"
#include "wsb_internal.h"

namespace wsb {

// ------------------------------------------------------------------------------ LayerNorm
// one warp per row; the row lives in registers between the statistics and the normalisation
constexpr int kLnMaxVec = 12;   // float4 per lane -> d <= 1536

__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ out_bf16,
                                                        float* __restrict__ out_f32, int rows, int d) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    if (row >= rows) return;
    const int nvec = d >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * d);
    float4 v[kLnMaxVec];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
        const int j = lane + 32 * i;
        if (j < nvec) {
            v[i] = xr[j];
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(sum) / d;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
        const int j = lane + 32 * i;
        if (j < nvec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            sq += (a * a + b * b) + (c * c + e * e);
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / d + 1e-5f);
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
        const int j = lane + 32 * i;
        if (j < nvec) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j);
            const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + j);
            const float o0 = (v[i].x - mean) * rstd * g.x + bt.x;
            const float o1 = (v[i].y - mean) * rstd * g.y + bt.y;
            const float o2 = (v[i].z - mean) * rstd * g.z + bt.z;
            const float o3 = (v[i].w - mean) * rstd * g.w + bt.w;
            if (out_bf16) {
                uint2 pk;
                pk.x = pack_bf16x2(o0, o1);
                pk.y = pack_bf16x2(o2, o3);
                reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * d)[j] = pk;
            }
            if (out_f32) reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * d)[j] = make_float4(o0, o1, o2, o3);
        }
    }
}

int layernorm_f32_to_bf16(const float* x, const float* gamma, const float* beta, __nv_bfloat16* out_bf16,
                          float* out_f32, int rows, int d, cudaStream_t stream) {
    WSB_REQUIRE(d % 4 == 0 && d <= kLnMaxVec * 128, "LayerNorm width must be a multiple of 4 and <= 1536");
    if (rows <= 0) return 0;
    const int rows_per_block = 8;
    WSB_CHECK_CUDA(launch_kernel(layernorm_kernel, dim3(ceil_div(rows, rows_per_block)), dim3(rows_per_block * 32), 0, stream, x,
                                 gamma, beta, out_bf16, out_f32, rows, d));
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------ split-K second phase
// One 128-thread CTA per output row (decode batches are a few hundred rows: a CTA per row keeps every
// load of the S partial planes independent and in flight at once).  The planes are summed in a fixed
// order (deterministic), then the fused epilogue runs: bias (+GELU) -> bf16, or bias + fp32 residual
// update (+ LayerNorm -> bf16 for the next GEMM), so a decoder layer needs no stand-alone LayerNorm.
constexpr int kRedThreads = 128;
constexpr int kRedMaxSplits = 16;

__global__ void __launch_bounds__(kRedThreads) splitk_reduce_bf16_kernel(const float* __restrict__ partial, int splits,
                                                                         long long split_stride, int N,
                                                                         const float* __restrict__ bias, int gelu,
                                                                         __nv_bfloat16* __restrict__ out,
                                                                         const unsigned char* __restrict__ row_skip) {
    const int row = blockIdx.x;
    const int nvec = N >> 2;
    pdl_wait();
    pdl_launch_dependents();
    // (the finished flag is requested together with the first planes: a dependent round trip less per launch)
    const bool skip = row_skip && row_skip[row];
    const float4* base = reinterpret_cast<const float4*>(partial + static_cast<long long>(row) * N);
    const long long sv = split_stride >> 2;
    for (int j = threadIdx.x; j < nvec; j += kRedThreads) {
        float4 t[kRedMaxSplits];
#pragma unroll
        for (int sp = 0; sp < kRedMaxSplits; ++sp)
            if (sp < splits) t[sp] = base[sp * sv + j];
        if (skip) return;
        float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int sp = 0; sp < kRedMaxSplits; ++sp)
            if (sp < splits) {
                acc.x += t[sp].x; acc.y += t[sp].y; acc.z += t[sp].z; acc.w += t[sp].w;
            }
        if (gelu) {
            acc.x = gelu_fast(acc.x); acc.y = gelu_fast(acc.y); acc.z = gelu_fast(acc.z); acc.w = gelu_fast(acc.w);
        }
        uint2 pk;
        pk.x = pack_bf16x2(acc.x, acc.y);
        pk.y = pack_bf16x2(acc.z, acc.w);
        reinterpret_cast<uint2*>(out + static_cast<long long>(row) * N)[j] = pk;
    }
}

constexpr int kRedLnVec = 3;   // float4 per thread -> N <= 1536

__global__ void __launch_bounds__(kRedThreads) splitk_reduce_resid_ln_kernel(const float* __restrict__ partial, int splits,
                                                                             long long split_stride, int N,
                                                                             const float* __restrict__ bias,
                                                                             float* __restrict__ x,
                                                                             const float* __restrict__ gamma,
                                                                             const float* __restrict__ beta,
                                                                             __nv_bfloat16* __restrict__ xn,
                                                                             const unsigned char* __restrict__ row_skip) {
    __shared__ float s_red[2][kRedThreads / 32];
    pdl_wait();
    pdl_launch_dependents();
    const int row = blockIdx.x;
    const bool skip = row_skip && row_skip[row];          // requested together with the planes below
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nvec = N >> 2;
    const float4* base = reinterpret_cast<const float4*>(partial + static_cast<long long>(row) * N);
    const long long sv = split_stride >> 2;
    float4* xr = reinterpret_cast<float4*>(x + static_cast<long long>(row) * N);
    float4 v[kRedLnVec];
    // the LayerNorm affine of this thread's elements: constants, fetched now instead of after the two reductions
    float4 gm[kRedLnVec], bt_[kRedLnVec];
    if (gamma != nullptr) {
#pragma unroll
        for (int i = 0; i < kRedLnVec; ++i) {
            const int j = tid + kRedThreads * i;
            if (j < nvec) {
                gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + j);
                bt_[i] = __ldg(reinterpret_cast<const float4*>(beta) + j);
            }
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kRedLnVec; ++i) {
        const int j = tid + kRedThreads * i;
        if (j < nvec) {
            float4 t[kRedMaxSplits];
#pragma unroll
            for (int sp = 0; sp < kRedMaxSplits; ++sp)
                if (sp < splits) t[sp] = base[sp * sv + j];
            float4 acc = xr[j];
            if (skip) return;
            if (bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + j);
                acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            }
#pragma unroll
            for (int sp = 0; sp < kRedMaxSplits; ++sp)
                if (sp < splits) {
                    acc.x += t[sp].x; acc.y += t[sp].y; acc.z += t[sp].z; acc.w += t[sp].w;
                }
            xr[j] = acc;
            v[i] = acc;
            sum += (acc.x + acc.y) + (acc.z + acc.w);
        }
    }
    if (gamma == nullptr) return;
    sum = warp_sum(sum);
    if (lane == 0) s_red[0][warp] = sum;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < kRedThreads / 32; ++i) tot += s_red[0][i];
    const float mean = tot / N;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < kRedLnVec; ++i) {
        const int j = tid + kRedThreads * i;
        if (j < nvec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            sq += (a * a + b * b) + (c * c + e * e);
        }
    }
    sq = warp_sum(sq);
    if (lane == 0) s_red[1][warp] = sq;
    __syncthreads();
    float tsq = 0.0f;
#pragma unroll
    for (int i = 0; i < kRedThreads / 32; ++i) tsq += s_red[1][i];
    const float rstd = rsqrtf(tsq / N + 1e-5f);
#pragma unroll
    for (int i = 0; i < kRedLnVec; ++i) {
        const int j = tid + kRedThreads * i;
        if (j < nvec) {
            const float4 g = gm[i];
            const float4 bt = bt_[i];
            uint2 pk;
            pk.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + bt.x, (v[i].y - mean) * rstd * g.y + bt.y);
            pk.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + bt.z, (v[i].w - mean) * rstd * g.w + bt.w);
            reinterpret_cast<uint2*>(xn + static_cast<long long>(row) * N)[j] = pk;
        }
    }
}

int splitk_reduce_bf16(const float* partial, int splits, int64_t split_stride, int M, int N, const float* bias, int gelu,
                       __nv_bfloat16* out, const unsigned char* row_skip, cudaStream_t stream) {
    WSB_REQUIRE(N % 4 == 0 && split_stride % 4 == 0 && splits <= kRedMaxSplits, "split-K reduce shape");
    if (M <= 0) return 0;
    WSB_CHECK_CUDA(launch_kernel(splitk_reduce_bf16_kernel, dim3(M), dim3(kRedThreads), 0, stream, partial, splits,
                                 static_cast<long long>(split_stride), N, bias, gelu, out, row_skip));
    count_launch();
    return 0;
}

int splitk_reduce_resid_ln(const float* partial, int splits, int64_t split_stride, int M, int N, const float* bias,
                           float* x, const float* gamma, const float* beta, __nv_bfloat16* xn,
                           const unsigned char* row_skip, cudaStream_t stream) {
    WSB_REQUIRE(N % 4 == 0 && N <= kRedLnVec * kRedThreads * 4 && split_stride % 4 == 0 && splits <= kRedMaxSplits,
                "split-K reduce shape (row width <= 1536)");
    if (M <= 0) return 0;
    WSB_CHECK_CUDA(launch_kernel(splitk_reduce_resid_ln_kernel, dim3(M), dim3(kRedThreads), 0, stream, partial, splits,
                                 static_cast<long long>(split_stride), N, bias, x, gamma, beta, xn, row_skip));
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------ conv1 + GELU
// out[b][1 + t][co] = gelu(bias[co] + sum_{ci,k} w[co][ci][k] * x[b][ci][t + k - 1]),  t in [0, n_cols)
// x: f32 [B][80][n_cols] (K1's output layout); wt: f32 [80*3][d] (pre-transposed, co contiguous);
// out: bf16, row 0 of every batch is the zero row conv2's left padding reads (written here).
constexpr int kC1Threads = 256;
constexpr int kC1TileT = 128;
constexpr int kC1TileC = 64;
constexpr int kC1In = 80;

__global__ void __launch_bounds__(kC1Threads) conv1_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                           int n_cols, int d, long long out_batch_stride) {
    extern __shared__ __align__(16) float c1_smem[];
    float* s_x = c1_smem;                               // [80][kC1TileT + 4]  (t0-1 .. t0+128)
    float* s_w = c1_smem + kC1In * (kC1TileT + 4);      // [240][64]
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * kC1TileT;
    const int co0 = blockIdx.y * kC1TileC;
    const int tid = threadIdx.x;
    const float* xb = x + static_cast<size_t>(b) * kC1In * n_cols;
    for (int i = tid; i < kC1In * (kC1TileT + 2); i += kC1Threads) {
        const int ci = i / (kC1TileT + 2), j = i - ci * (kC1TileT + 2);
        const int t = t0 - 1 + j;
        s_x[ci * (kC1TileT + 4) + j] = (t >= 0 && t < n_cols) ? __ldg(xb + static_cast<size_t>(ci) * n_cols + t) : 0.0f;
    }
    for (int i = tid; i < kC1In * 3 * (kC1TileC / 4); i += kC1Threads) {
        const int kk = i / (kC1TileC / 4), c4 = i - kk * (kC1TileC / 4);
        reinterpret_cast<float4*>(s_w)[kk * (kC1TileC / 4) + c4] =
            __ldg(reinterpret_cast<const float4*>(wt + static_cast<size_t>(kk) * d + co0) + c4);
    }
    __syncthreads();
    const int cx = tid & 15, ty = tid >> 4;             // 16 x 16 threads: 4 channels x 8 time steps each
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
    for (int ci = 0; ci < kC1In; ++ci) {
        float xv[10];
        const float* xs = s_x + ci * (kC1TileT + 4) + ty * 8;
#pragma unroll
        for (int i = 0; i < 10; ++i) xv[i] = xs[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 w4 = reinterpret_cast<const float4*>(s_w)[(ci * 3 + k) * (kC1TileC / 4) + cx];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][0] = fmaf(w4.x, xv[i + k], acc[i][0]);
                acc[i][1] = fmaf(w4.y, xv[i + k], acc[i][1]);
                acc[i][2] = fmaf(w4.z, xv[i + k], acc[i][2]);
                acc[i][3] = fmaf(w4.w, xv[i + k], acc[i][3]);
            }
        }
    }
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bias + co0) + cx);
    __nv_bfloat16* ob = out + static_cast<size_t>(b) * out_batch_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + ty * 8 + i;
        if (t < n_cols) {
            uint2 pk;
            pk.x = pack_bf16x2(gelu_erf(acc[i][0] + bs.x), gelu_erf(acc[i][1] + bs.y));
            pk.y = pack_bf16x2(gelu_erf(acc[i][2] + bs.z), gelu_erf(acc[i][3] + bs.w));
            *reinterpret_cast<uint2*>(ob + static_cast<size_t>(1 + t) * d + co0 + cx * 4) = pk;
        }
    }
    if (blockIdx.x == 0 && ty == 0)                     // zero row in front of every batch
        *reinterpret_cast<uint2*>(ob + co0 + cx * 4) = make_uint2(0u, 0u);
}

int conv1_gelu(const float* feats, const float* wt, const float* b, __nv_bfloat16* out, int B, int n_cols, int d,
               int64_t out_batch_stride, cudaStream_t stream) {
    WSB_REQUIRE(d % kC1TileC == 0, "d_model must be a multiple of 64");
    if (B <= 0) return 0;
    static PerDeviceOnce once;
    int dev = 0;
    const size_t smem = sizeof(float) * (kC1In * (kC1TileT + 4) + kC1In * 3 * kC1TileC);
    if (once.need(&dev)) {
        WSB_CHECK_CUDA(cudaFuncSetAttribute(conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        once.mark(dev);
    }
    dim3 grid(ceil_div(n_cols, kC1TileT), d / kC1TileC, B);
    conv1_kernel<<<grid, kC1Threads, smem, stream>>>(feats, wt, b, out, n_cols, d, out_batch_stride);
    WSB_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------ features -> time-major bf16 (conv1 GEMM operand)
__global__ void __launch_bounds__(256) features_tm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int n_cols) {
    // every feature is written as a bf16 pair hi = bf16(x), lo = bf16(x - hi): the GEMM sums both against the same weight, so the
    // operand keeps ~16 mantissa bits (the fp32 conv1 it replaces saw the features unrounded)
    __shared__ float tile[80][33];
    const int b = blockIdx.y, t0 = blockIdx.x * 32, tid = threadIdx.x;
    const float* xb = x + static_cast<size_t>(b) * kC1In * n_cols;
    for (int i = tid; i < kC1In * 32; i += 256) {
        const int ci = i >> 5, j = i & 31;
        tile[ci][j] = (t0 + j < n_cols) ? __ldg(xb + static_cast<size_t>(ci) * n_cols + t0 + j) : 0.0f;
    }
    __syncthreads();
    constexpr int kRow = 2 * kC1In;
    __nv_bfloat16* ob = out + static_cast<size_t>(b) * (n_cols + 3) * kRow;
    for (int i = tid; i < 32 * kC1In; i += 256) {
        const int j = i / kC1In, ci = i - j * kC1In;
        if (t0 + j < n_cols) {
            const float v = tile[ci][j];
            const __nv_bfloat16 hi = __float2bfloat16(v);
            __nv_bfloat16* o = ob + static_cast<size_t>(1 + t0 + j) * kRow;
            o[ci] = hi;
            o[kC1In + ci] = __float2bfloat16(v - __bfloat162float(hi));
        }
    }
    if (blockIdx.x == 0 && tid < kRow) {                 // the zero rows: conv padding (row 0, row n_cols + 1) and the K tail (row n_cols + 2)
        ob[tid] = __float2bfloat16(0.0f);
        ob[static_cast<size_t>(n_cols + 1) * kRow + tid] = __float2bfloat16(0.0f);
        ob[static_cast<size_t>(n_cols + 2) * kRow + tid] = __float2bfloat16(0.0f);
    }
}

int features_time_major_bf16(const float* feats, __nv_bfloat16* out, int B, int n_cols, cudaStream_t stream) {
    if (B <= 0) return 0;
    WSB_CHECK_CUDA(launch_kernel(features_tm_kernel, dim3(ceil_div(n_cols, 32), B), dim3(256), 0, stream, feats, out, n_cols));
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------ decoder embedding
__global__ void embed_kernel(const int* __restrict__ tokens, const int* __restrict__ step_ptr, int pos_offset,
                             const __nv_bfloat16* __restrict__ emb, const float* __restrict__ pos_emb,
                             float* __restrict__ x, int d) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x;
    const int tok = tokens[b];
    const int pos = pos_offset + (step_ptr ? *step_ptr : 0);
    const __nv_bfloat16* e = emb + static_cast<size_t>(tok) * d;
    const float* pe = pos_emb + static_cast<size_t>(pos) * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[static_cast<size_t>(b) * d + i] = __bfloat162float(e[i]) + pe[i];
}

int embed_tokens_step(const int* tokens, const int* step_ptr, int pos_offset, const __nv_bfloat16* emb,
                      const float* pos_emb, float* x, int B, int d, cudaStream_t stream) {
    if (B <= 0) return 0;
    WSB_CHECK_CUDA(launch_kernel(embed_kernel, dim3(B), dim3(256), 0, stream, tokens, step_ptr, pos_offset, emb, pos_emb, x, d));
    count_launch();
    return 0;
}

}  // namespace wsb
