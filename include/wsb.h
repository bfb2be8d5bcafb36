/* libwsb -- C ABI of the B200-native WhisperSeg segmentation hot path.
 *
 * The reference (nianlonggu/WhisperSeg) is pure Python and has NO plugin / operator / FFI layer: its
 * hot path sits behind the Python methods of `model.py`.  This header is therefore the FFI a
 * maintainer would bind *from that Python file* (ctypes).  Every entry point
 * names the reference interface it replaces.  Conventions: plain pointers and sizes, no torch
 * types; device pointers are raw CUDA device addresses; `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream); every function returns 0 on success and a non-zero
 * status otherwise, with a human-readable message available from wsb_last_error().
 * There is no CPU fallback: every compute entry point requires an sm_100 device.
 */
#ifndef WSB_H_
#define WSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSB_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------------ */
int wsb_abi_version(void);
const char* wsb_last_error(void);
/* number of CUDA kernels launched by this library since the last call with reset != 0 */
long long wsb_launch_count(int reset);
/* Chunk pipelining (segmenter.py: the encoder of chunk i+1 runs while chunk i decodes; the reference runs them
 * strictly one after the other, model.py:653-668): the persistent tcgen05 GEMMs launched BY THE CALLING THREAD from
 * now on occupy at most (#SMs - n_sms) SMs, leaving the rest to the latency-bound decode kernels of other streams.
 * Returns the previous value; 0 restores full-width launches.                                                    */
int wsb_set_sm_reserve(int n_sms);

/* ---- K1: log-mel front-end ------------------------------------------------------------------
 * Replaces SegmenterBase.get_sliced_audios_features's per-window feature extraction
 * (reference model.py:146-165 -> HF WhisperFeatureExtractor, called at model.py:152) and the
 * WhisperSegFeatureExtractor configuration (reference audio_utils.py:45-76).                   */
typedef struct wsb_logmel_plan wsb_logmel_plan;

/* mel_filters: host float32 [n_fft/2+1][80] (the slaney bank for [min_frequency, sr//2]);
 * n_cols = total_spec_columns (1000).  clip_len = int(n_cols * spec_time_step * sr).           */
int wsb_logmel_plan_create(int n_fft, int hop, int clip_len, int n_cols, const float* mel_filters,
                           int n_freq, wsb_logmel_plan** plan);
void wsb_logmel_plan_destroy(wsb_logmel_plan* plan);
/* audio_dev: device float32 samples (16-byte aligned).  windows_dev: device int64 [n_windows][3] =
 * {start, lo, hi}: window w covers audio[start .. start+clip_len) and a sample index a is read
 * only if lo <= a < hi (zero otherwise: trial left-padding, tail zero-padding, file boundaries).
 * features_dev: device float32 [n_windows][80][n_cols].                                         */
int wsb_logmel_run(const wsb_logmel_plan* plan, const float* audio_dev, const int64_t* windows_dev,
                   int n_windows, float* features_dev, void* stream);

/* ---- model ------------------------------------------------------------------------------------
 * Replaces WhisperForConditionalGeneration as used by WhisperSegmenter.__init__ /
 * generate_segment_text_core (reference model.py:626-676) and ctranslate2.models.Whisper as used
 * by WhisperSegmenterFast (model.py:679-746).                                                   */
typedef struct wsb_model wsb_model;

typedef struct wsb_model_config {
    int d_model;            /* 384 / 512 / 768 / 1280 ...; head_dim is 64 */
    int n_heads;
    int n_layers;           /* encoder_layers == decoder_layers */
    int ffn_dim;
    int vocab_size;         /* 51865 */
    int n_mels;             /* 80 */
    int n_cols;             /* total_spec_columns: 1000 -> 500 encoder positions */
    int max_target_positions; /* 448 */
    int max_batch;          /* windows processed per encode/generate call */
} wsb_model_config;

/* names[i] / tensors_dev[i]: prepared device tensors (see whisperseg_b200/weights.py for the list,
 * dtypes and layouts).  The model keeps the pointers; the caller keeps the memory alive.        */
int wsb_model_create(const wsb_model_config* cfg, const char* const* names, const void* const* tensors_dev,
                     int n_tensors, wsb_model** model);
void wsb_model_destroy(wsb_model* model);
size_t wsb_model_workspace_bytes(const wsb_model* model);
/* Workspace a model of this configuration would allocate (no device needed): lets the host pick max_batch from the
 * free device memory when the caller -- e.g. the reference's unchanged scripts/segment.py -- does not give one. */
size_t wsb_workspace_bytes_for(const wsb_model_config* cfg);
/* 1 once the folded-LayerNorm guard has fired for this model: at <= 64 decode rows the LayerNorm is folded into the
 * projection that consumes it (the kernel multiplies bf16(x), not bf16(LN(x))), which is only as accurate as the
 * exact form while a row's |mean| stays below ~2 standard deviations; when a live row violates that, the engine
 * switches this model to the exact on-the-fly LayerNorm for good (HF nn.LayerNorm semantics, modeling_whisper.py:417-506). */
int wsb_model_fold_fallback(const wsb_model* model);
/* Diagnostics of the persistent decode kernel (csrc/mega.cu): out_host == NULL arms the trace and returns its length in
 * words; otherwise copies the %globaltimer stamps (ns) CTA 0 took during the LAST decoder position into out_host:
 * word 0 = start, word 2k-1 = phase k's work done, word 2k = grid barrier k passed.  Returns the words copied (> 0);
 * unlike the other entry points a return value <= 0 is the error.                                                  */
int wsb_mega_trace(wsb_model* model, unsigned long long* out_host, int n);

/* Encoder (conv stem + n_layers pre-LN blocks + final LN); replaces HF WhisperEncoder.forward
 * reached from model.generate (reference model.py:655).  features_dev: float32 [batch][80][n_cols].
 * hidden_f32_dev: optional (may be NULL) float32 [batch][n_cols/2][d_model] copy of the output;
 * the bf16 output stays inside the model for wsb_generate.                                       */
int wsb_encode(wsb_model* model, const float* features_dev, int batch, float* hidden_f32_dev, void* stream);

/* Greedy generation over the windows of the last wsb_encode call; replaces model.generate(...,
 * num_beams=1) at reference model.py:655-666 / 723-727.  prompt: host int32[prompt_len]
 * ([<|startoftranscript|>, <|en|>, <|notimestamps|>]); max_length counts the prompt tokens.
 * tokens_dev: device int32 [batch][max_length - prompt_len], finished rows padded with pad_id.
 * forced_dev: NULL, or device int32 [batch][max_length] decoder inputs for teacher forcing (the
 * per-position arg-max is still what is written to tokens_dev).  n_steps (host, may be NULL)
 * receives the number of generated positions actually computed.  flags: bit0 = replay the decode
 * step from a CUDA graph; bit1 = launch decode kernels with programmatic dependent launch; bit2 =
 * disable batch compaction (gathering the still-active rows into a smaller batch as rows finish); bit3 =
 * keep the tcgen05 split-K linear layers for batches of <= 64 rows (default: fused LN + mma.sync kernels). */
int wsb_generate(wsb_model* model, int batch, const int32_t* prompt, int prompt_len, int eos_id, int pad_id,
                 int max_length, const int32_t* forced_dev, int32_t* tokens_dev, int* n_steps, int flags,
                 void* stream);

/* Beam search over the windows of the last wsb_encode call; replaces model.generate(..., num_beams=4,
 * length_penalty=1.0) -- the reference's DEFAULT decode mode (reference model.py:409, 614, 662, 724) --
 * with HF semantics (transformers generation/utils.py `_beam_search`, early_stopping=False): fp32
 * log-softmax over the full vocabulary, then the suppression masks; 2*num_beams continuations per step;
 * finished hypotheses scored sum_logprob / generated_len**length_penalty; the best one is returned.
 * batch*num_beams <= max_batch, num_beams in [1,4].  tokens_dev: device int32 [batch][max_length -
 * prompt_len] (EOS included, then pad_id); scores_dev: optional device float32 [batch] (score of the
 * returned hypothesis).  flags: bit0 = replay the decode step from a CUDA graph.                      */
int wsb_generate_beam(wsb_model* model, int batch, int num_beams, const int32_t* prompt, int prompt_len, int eos_id,
                      int pad_id, int max_length, float length_penalty, int32_t* tokens_dev, float* scores_dev,
                      int* n_steps, int flags, void* stream);

/* The beam bookkeeping kernels alone, driven by caller-provided logits (parity tests): logits_dev float32
 * [n_steps][batch*num_beams][vocab], suppress_dev float32 [vocab] additive mask.  Outputs as
 * wsb_generate_beam plus, per step, the parent row of every surviving beam (parents_dev int32
 * [n_steps][batch*num_beams]) and its token (next_tokens_dev, same shape); both optional.           */
int wsb_beam_selftest(int batch, int num_beams, int vocab, int n_steps, const float* logits_dev, const float* suppress_dev,
                      const int32_t* prompt, int prompt_len, int eos_id, int pad_id, int max_length, float length_penalty,
                      int32_t* tokens_dev, float* scores_dev, int32_t* parents_dev, int32_t* next_tokens_dev, void* stream);

/* ---- CUDA-event profiling of kernel classes (bench.py's roofline numbers) -----------------------
 * While enabled, every eagerly launched kernel is bracketed by CUDA events on its own stream.
 * categories: 0 conv1, 1 encoder GEMMs (conv2, qkv, out, fc1, fc2), 2 encoder attention,
 * 3 encoder LayerNorm, 4 cross-K/V GEMM, 5 decoder GEMMs, 6 logits+arg-max GEMM, 7 decode
 * self-attention, 8 decode cross-attention, 9 decoder LayerNorm, 10 misc, 11 CUDA-graph replays of a whole decoder
 * position (bracketed as one unit: `work` counts positions), 12 batch compaction.
 * wsb_profile_read: call after synchronising the stream; `work` = accumulated algorithmic FLOPs
 * (GEMMs, attention) or bytes (LayerNorm, decode cross-attention) of the bracketed launches.      */
int wsb_profile_enable(int enable);
int wsb_profile_read(int category, double* ms, long long* launches, double* work);

/* ---- individual kernels, exported for parity tests and profiling ------------------------------ */
/* C = act(A W^T + bias) (+ resid): A bf16 [M][K], W bf16 [N][K], fp32 accumulate.
 * out_f32 != 0: C float32 [M][N] (+ optional float32 residual [M][N]); else C bf16 [M][N].      */
int wsb_gemm_bf16(const void* a_dev, const void* w_dev, int M, int N, int K, const float* bias_dev, int gelu,
                  const float* resid_dev, void* c_dev, int out_f32, int block_n, void* stream);
int wsb_layernorm(const float* x_dev, const float* gamma_dev, const float* beta_dev, void* out_bf16_dev,
                  float* out_f32_dev, int rows, int d, void* stream);
/* Skinny linear layer for decode batches of at most 64 rows (csrc/gemv.cu): out = epilogue(in * W^T + bias) in
 * one launch.  Input: fp32 rows x_f32_dev [M][K] with the LayerNorm (gamma, beta) fused in, or (x NULL) bf16
 * activations a_bf16_dev [M][K].  w_dev bf16 [N][K].  out_mode 0: float32 [M][N]; 1: bf16 [M][N] = GELU(.);
 * 2: float32 [M][N] += (in-place residual update).  Folded-LayerNorm form (x given, beta_dev NULL): w_dev =
 * bf16(W o gamma), gamma_dev = c1 [N] (row sums of w_dev), bias_dev = c2 = b + W beta; the kernel reads bf16(x) and
 * applies rstd (acc - mean c1) + c2.  Replaces nn.Linear (+ the preceding nn.LayerNorm) of
 * HF WhisperDecoderLayer (modeling_whisper.py:417-506) at small batch.                                       */
int wsb_gemv16(const float* x_f32_dev, const float* gamma_dev, const float* beta_dev, const void* a_bf16_dev,
               const void* w_dev, const float* bias_dev, int M, int N, int K, int out_mode, void* out_dev, void* stream);

/* Skinny linear layer for decode batches of 65..256 rows (csrc/skinny.cu): one launch, split-K across a thread-block
 * cluster (splits = 0: chosen; else 1, 2, 4 or 8), partial tiles reduced through distributed shared memory.
 * Folded-LayerNorm form (x_f32_dev [M][K] given): w_dev = bf16(W o gamma) [N][K], c1_dev [N] = its row sums, bias_dev =
 * c2 = b + W beta; out_mode 0: float32 [M][N], 1: bf16 [M][N] = GELU(.).  Residual form (x NULL, a_bf16_dev [M][K]):
 * out_mode 2, out_dev float32 [M][N] += a W^T + bias, and optionally xb_out_dev bf16 [M][N] = the updated rows,
 * stats_out_dev float32 [N / 128][M][2] = per-tile (sum, sum of squares) of every updated row.  N % 128 == 0,
 * K % 64 == 0.  Replaces nn.Linear (+ the preceding nn.LayerNorm) of HF WhisperDecoderLayer
 * (modeling_whisper.py:417-506) at medium batch.                                                              */
int wsb_skinny_linear(const float* x_f32_dev, const float* c1_dev, const void* a_bf16_dev, const void* w_dev,
                      const float* bias_dev, int M, int N, int K, int out_mode, int splits, void* out_dev,
                      void* xb_out_dev, float* stats_out_dev, void* stream);

/* Diagnostics: steady-state microseconds per wsb_gemv16-style launch (weights rotating through `weight_copies`
 * buffers so that they stream from HBM).  mode 0: LayerNorm -> fp32, 1: LayerNorm -> GELU bf16, 2: bf16 -> residual,
 * 3 / 4: folded LayerNorm -> fp32 / GELU bf16. */
int wsb_gemv16_bench(int M, int N, int K, int mode, int iters, int weight_copies, float* us_per_launch);

/* qkv bf16 [batch*T][3*n_heads*64] (q pre-scaled) -> out bf16 [batch*T][n_heads*64] */
int wsb_encoder_attention(const void* qkv_dev, void* out_dev, int batch, int T, int n_heads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WSB_H_ */
