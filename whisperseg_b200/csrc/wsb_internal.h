// Internal C++ interfaces between the translation units of libwsb.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stddef.h>
#include <stdint.h>

namespace wsb {

// ------------------------------------------------------------------ K1 log-mel (logmel.cu)
struct LogmelPlan;
int logmel_plan_create(int n_fft, int hop, int clip_len, int n_cols, const float* mel_filters_host, int n_freq,
                       LogmelPlan** out);
void logmel_plan_destroy(LogmelPlan* pl);
int logmel_run(const LogmelPlan* pl, const float* audio_dev, const long long* win_dev, int n_win, float* out_dev,
               cudaStream_t stream);
size_t logmel_plan_smem(const LogmelPlan* pl);
int logmel_plan_cluster(const LogmelPlan* pl);

// ------------------------------------------------------------------ K3 GEMM (gemm.cu)
// C[M,N] = A[M,K] (bf16, K contiguous) * W[N,K]^T (bf16, K contiguous), fp32 accumulation in TMEM.
enum GemmOut : int {
    GEMM_OUT_BF16 = 0,        // out bf16 [M, ldc]
    GEMM_OUT_F32 = 1,         // out f32  [M, ldc]
    GEMM_OUT_HEADMAJOR = 2,   // out bf16 [M/rows_per_batch][N/64][rows_per_batch][64]
    GEMM_OUT_ARGMAX = 3,      // out: per (row, n-tile) partial {max, argmax} -- never materialises [M,N]
};
enum GemmAct : int { GEMM_ACT_NONE = 0, GEMM_ACT_GELU = 1 };

struct GemmArgs {
    const __nv_bfloat16* A = nullptr;   // activations
    int64_t lda = 0;                    // elements between consecutive rows of A
    int a_rows_per_batch = 0;           // 0: flat [M, K]; else A is [batches][a_rows_per_batch] rows with
    int64_t a_batch_stride = 0;         //    a_batch_stride elements between batches (M = batches * rows)
    const __nv_bfloat16* W = nullptr;   // [N, K] row-major (ldw = K)
    int M = 0, N = 0, K = 0;
    const float* bias = nullptr;        // [N] or null
    const float* bias2 = nullptr;       // [N] or null (second additive vector, e.g. begin-suppress mask)
    int act = GEMM_ACT_NONE;
    const float* resid = nullptr;       // f32 [M, ldr] added after activation (may alias out)
    int64_t ldr = 0;
    const float* rowvec = nullptr;      // f32 [rows_per_batch, N] added per (row % rows_per_batch) (pos-emb)
    int out_mode = GEMM_OUT_BF16;
    void* out = nullptr;
    int64_t ldc = 0;
    int rows_per_batch = 0;             // for HEADMAJOR / rowvec
    float* argmax_val = nullptr;        // [M, n_tiles]
    int* argmax_idx = nullptr;          // [M, n_tiles]
    int block_n = 0;                    // 0 = choose
    int splits = 1;                     // split-K factor (GEMM_OUT_F32 without bias only): partial plane s is
    int64_t split_stride = 0;           //   written at out + s * split_stride; see splitk_reduce_* below
    const unsigned char* row_skip = nullptr;   // optional [M] flags: flagged rows are not stored (GEMM_OUT_F32)
    int c_batch_pad = 0;                // GEMM_OUT_BF16 with a_rows_per_batch > 0: C has this many extra rows between batches
    const int* n_tile_list = nullptr;   // optional device list of the n-tiles (of width block_n) to compute; the
    int n_tile_count = 0;               //   arg-max partials are then [M][n_tile_count]
};
// effective number of partial planes gemm_bf16 will write for (K, splits)
int gemm_effective_splits(int K, int splits);
// pick (block_n, splits) for a skinny GEMM so that ~all SMs stream distinct weight bytes
void gemm_pick_skinny(int M, int N, int K, int* block_n, int* splits);
int gemm_bf16(const GemmArgs& a, cudaStream_t stream);
int gemm_n_tiles(int N, int block_n);
int gemm_pick_block_n(int M, int N);

// ------------------------------------------------------------------ small kernels (elementwise.cu)
int layernorm_f32_to_bf16(const float* x, const float* gamma, const float* beta, __nv_bfloat16* out_bf16,
                          float* out_f32, int rows, int d, cudaStream_t stream);
int conv1_gelu(const float* feats, const float* w /*[d][80][3]*/, const float* b, __nv_bfloat16* out, int B,
               int n_cols, int d, int64_t out_batch_stride, cudaStream_t stream);
// f32 features [B][80][n_cols] -> bf16 time-major [B][n_cols + 3][160] (80 hi parts, then 80 lo parts = bf16(x - hi)) with a
// zero row in front and two behind: the (k = 3, pad = 1) conv1 window of frame t is then the contiguous span of 480 elements
// starting at row t (im2col-free GEMM operand)
int features_time_major_bf16(const float* feats, __nv_bfloat16* out, int B, int n_cols, cudaStream_t stream);
// split-K second phase: out = epilogue(sum_s partial[s]) -- deterministic summation order.
//   bf16 variant : out_bf16[M][N] = act(sum + bias)
//   resid+LN     : x[M][N] += sum + bias (fp32, in place); if gamma: xn_bf16 = LayerNorm(x) (N <= 1536)
int splitk_reduce_bf16(const float* partial, int splits, int64_t split_stride, int M, int N, const float* bias, int gelu,
                       __nv_bfloat16* out, const unsigned char* row_skip, cudaStream_t stream);
int splitk_reduce_resid_ln(const float* partial, int splits, int64_t split_stride, int M, int N, const float* bias,
                           float* x, const float* gamma, const float* beta, __nv_bfloat16* xn,
                           const unsigned char* row_skip, cudaStream_t stream);
int embed_tokens(const int* tokens, const int* positions, const __nv_bfloat16* emb, const float* pos_emb, float* x,
                 int B, int d, cudaStream_t stream);

// ------------------------------------------------------------------ K5c skinny linear for <= 64 rows (gemv.cu)
// out = epilogue(in[M,K] * W[N,K]^T + bias), M <= 64, one launch, no split-K.  MP = 16, 32 or 64 = M rounded up to
// 1, 2 or 4 m-tiles.  Input: either fp32 rows `x` with
// the consumer's LayerNorm (gamma, beta) fused in -- `stats` = [stats_parts][MP][2] partial (sum, sum of squares)
// of every row, left by the producer of x -- or bf16 activations `a`.  Output: exactly one of fp32 `out_f32`,
// bf16 `out_bf16_gelu` (GELU applied) or the in-place fp32 residual update `resid` (+=), which also writes the
// updated rows' partial statistics to `stats_out` [N / 8][MP][2] (one partial per 8-feature n-tile) when given.
struct Gemv16Args {
    const float* x = nullptr;
    const float* stats = nullptr;
    int stats_parts = 0;
    float* stats_out = nullptr;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    const __nv_bfloat16* a = nullptr;
    const float* c1 = nullptr;           // folded LayerNorm: a = bf16(x), W = bf16(W o gamma), c1 = row sums of W, bias = b + W beta
    __nv_bfloat16* xb_out = nullptr;     // residual mode: also write the updated rows as bf16
    const __nv_bfloat16* W = nullptr;
    const float* bias = nullptr;
    float* out_f32 = nullptr;
    __nv_bfloat16* out_bf16_gelu = nullptr;
    float* resid = nullptr;
    const unsigned char* row_skip = nullptr;
    int* fold_flag = nullptr;            // folded LayerNorm guard: set to 1 when a live row has |mean| > 2 std
    int M = 0, N = 0, K = 0;
};
int gemv16(const Gemv16Args& a, cudaStream_t stream);
int gemv16_max_rows();
int gemv16_parts(int N, int K);                                 // CTAs (= statistics partials) of a launch with N outputs
int row_stats16(const float* x, int M, int K, float* stats, cudaStream_t stream,
                __nv_bfloat16* xb = nullptr);            // exact statistics, 1 part (+ optional bf16 copy of x)

// ------------------------------------------------------------------ K5e persistent decode position for <= 64 rows (mega.cu)
// One launch runs embedding + every decoder layer of one position (grid barriers between phases, weight tiles of the
// next two linear-layer jobs in flight).  `layers_dev`: device table built with mega_fill_layer (per layer the six
// linear layers qkv, self-out, cross-q, cross-out, fc1, fc2: W / bias / c1, the LayerNorm-consuming ones in folded form).
// Leaves the updated residual stream in dx (+ bf16 copy, row statistics); the caller applies the final LayerNorm.
struct MegaArgs {
    const void* layers_dev = nullptr;
    int L = 0, d = 0, F = 0, H = 0, T = 0, tmax = 0, B = 0;
    const int* next_token = nullptr;
    const int* step_ptr = nullptr;
    const __nv_bfloat16* emb = nullptr;
    const float* pos_emb = nullptr;
    float* dx = nullptr;
    __nv_bfloat16* dxn = nullptr;
    float* stats = nullptr;
    float* proj = nullptr;
    __nv_bfloat16* datt = nullptr;
    __nv_bfloat16* dff = nullptr;
    __nv_bfloat16* k_cache = nullptr;
    __nv_bfloat16* v_cache = nullptr;
    const __nv_bfloat16* cross_kv = nullptr;
    const unsigned char* finished = nullptr;
    int kv_div = 1;
    unsigned int* sync = nullptr;        // device, 64 words, zero before the first launch: arrivals, exits, watchdog flag, ..., generation
    int* fold_flag = nullptr;
    unsigned long long* trace = nullptr; // diagnostics: device buffer of 2 * (2 + 10 L) timestamps (wsb_mega_trace)
};
bool mega_supported(int d, int F);
int decode_layers_mega(const MegaArgs& a, cudaStream_t stream);
size_t mega_layer_table_bytes(int n_layers);
void mega_fill_layer(void* host_table, int layer, const void* const W[6], const float* const bias[6], const float* const c1[6]);

// ------------------------------------------------------------------ K5d skinny linear for 65..256 rows (skinny.cu)
// One launch per linear layer: split-K across a thread-block cluster, partial tiles reduced through DSMEM, LayerNorm
// folded into the projection (A = bf16 rows of the residual stream, W = bf16(W o gamma), c1 = row sums of W, bias = c2,
// `stats` = [stats_parts][stats_ld][2] partial (sum, sum of squares) per row) -> fp32 `out_f32` | GELU bf16
// `out_bf16_gelu`; or the in-place residual update `resid` += A W^T + bias (A = bf16 activations), which also writes
// the updated rows as bf16 (`xb_out`) and this launch's per-tile row statistics `stats_out` [N / 128][stats_out_ld][2].
struct SkinnyArgs {
    const __nv_bfloat16* A = nullptr;
    int64_t lda = 0;
    const __nv_bfloat16* W = nullptr;    // [N][K]
    int M = 0, N = 0, K = 0;
    const float* bias = nullptr;
    const float* c1 = nullptr;
    const float* stats = nullptr;
    int stats_parts = 0, stats_ld = 0;
    float* out_f32 = nullptr;
    __nv_bfloat16* out_bf16_gelu = nullptr;
    float* resid = nullptr;
    __nv_bfloat16* xb_out = nullptr;
    float* stats_out = nullptr;
    int stats_out_ld = 0;
    const unsigned char* row_skip = nullptr;
    int splits = 0;                      // 0 = choose (1, 2, 4 or 8 CTAs per cluster along K)
};
bool skinny_cluster_supported(int M, int N, int K);
int skinny_cluster_splits(int M, int N, int K);
int skinny_cluster_linear(const SkinnyArgs& a, cudaStream_t stream);
int row_stats_any(const float* x, int M, int K, float* stats, __nv_bfloat16* xb, cudaStream_t stream);   // gemv.cu: any M

// ------------------------------------------------------------------ K4 encoder attention (attention.cu)
int encoder_attention(const __nv_bfloat16* qkv /*[B*T, 3d]*/, __nv_bfloat16* out /*[B*T, d]*/, int B, int T,
                      int n_heads, cudaStream_t stream);

// ------------------------------------------------------------------ K5 decode kernels (decode.cu)
struct DecodeAttnArgs;

}  // namespace wsb
