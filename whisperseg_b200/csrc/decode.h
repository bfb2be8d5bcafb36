// Declarations of the decode-step kernels (decode.cu) and the embedding helper (elementwise.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace wsb {

// fp32 split-K partial planes of the projection feeding an attention kernel (fused second phase)
struct SplitkInput {
    const float* planes;
    int splits;
    long long split_stride;
    const float* bias;
};
// `anc` (beam search): ancestry tables [2][B][anc_ld] -- keys at position j < pos of row b are read from the
// cache block of row anc[pos & 1][b][j]; null = every row reads its own block.
int decode_self_attention(const __nv_bfloat16* qkv, const SplitkInput* part, int d, __nv_bfloat16* k_cache,
                          __nv_bfloat16* v_cache, int t_max, const int* step_ptr, int pos_offset,
                          const unsigned char* finished, __nv_bfloat16* out, int B, int n_heads, cudaStream_t stream,
                          const int* anc = nullptr, int anc_ld = 0);
// `kv_div` (beam search): consecutive rows sharing one window's cross-attention K/V block.
int decode_cross_attention(const __nv_bfloat16* q, const SplitkInput* part, int d, const __nv_bfloat16* cross_kv, int layer,
                           int n_layers, int T, const unsigned char* finished, __nv_bfloat16* out, int B, int n_heads,
                           cudaStream_t stream, int kv_div = 1);
int argmax_finalize(const float* val, const int* idx, int n_tiles, int* tokens_out, int max_new, int out_offset,
                    int* next_token, const int* forced, int forced_ld, unsigned char* finished, int* step_ptr,
                    int* n_active, int eos_id, int pad_id, int B, const int* row_map, cudaStream_t stream);
// gather the still-active rows of a decode batch into a smaller dense batch (see decode.cu)
struct CompactArgs {
    const unsigned char* fin_src;
    int b_src;
    const int* map_src;             // slot -> original window of the source batch (null = identity)
    const int* tok_src;
    int b_dst;
    int* active_idx;                // [b_dst] scratch: source row of every destination slot (-1 = padding)
    int* map_dst;
    int* tok_dst;
    unsigned char* fin_dst;
    const __nv_bfloat16 *k_src, *v_src, *cross_src;
    __nv_bfloat16 *k_dst, *v_dst, *cross_dst;
    const int* step_ptr;
    int n_heads, n_layers, t_max;
    long long cross_row_elems;
};
int compact_decode_state(const CompactArgs& a, cudaStream_t stream);
int prefill_advance(int* next_token, const int* forced, int forced_ld, const int* prompt_dev, int* step_ptr, int B,
                    cudaStream_t stream);
int embed_tokens_step(const int* tokens, const int* step_ptr, int pos_offset, const __nv_bfloat16* emb,
                      const float* pos_emb, float* x, int B, int d, cudaStream_t stream);

// ---- prompt prefill: the P prompt positions of all B rows in ONE pass (virtual row p * B + b = position p of row b).
// x[p * B + b] = emb[token] + pos_emb[p]; token = forced[b][p] if forced else prompt_dev[p]
int embed_prefill(const int* prompt_dev, const int* forced, int forced_ld, const __nv_bfloat16* emb, const float* pos_emb,
                  float* x, int B, int P, int d, cudaStream_t stream);
// causal self-attention over the P prompt positions of every (row, head): reads q, k, v of the virtual rows from the fp32
// projection planes (+ bias), writes the K/V cache positions 0..P-1 and the attention outputs of all P positions
int prefill_self_attention(const SplitkInput* part, int d, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int t_max,
                           __nv_bfloat16* out, int B, int P, int n_heads, cudaStream_t stream);
// cross-attention of the P prompt positions of a row from ONE pass over the row's K/V block (the per-position path streams
// the 128 KB block once per position)
int prefill_cross_attention(const SplitkInput* part, int d, const __nv_bfloat16* cross_kv, int layer, int n_layers, int T,
                            __nv_bfloat16* out, int B, int P, int n_heads, cudaStream_t stream);

// ---- beam search (beam.cu): device-resident state of `B` windows x `nb` beams
struct BeamState {
    int B, nb, K;                       // windows, beams per window, continuations kept per step (2*nb)
    int V, max_length, prompt_len, eos_id, pad_id, seq_ld;
    long long ldv;                      // row stride of `logits`
    const float* logits;                // raw fp32 logits [B*nb][ldv] of the current position
    const float* suppress;              // additive masks [V] (0 / -inf), applied after the log-softmax
    const float* begin_suppress;        // first generated position only (nullable)
    float* cand_val;                    // [B*nb][8] per-row top continuations (score, token)
    int* cand_idx;
    int* run_seq;                       // [2][B*nb][seq_ld] token histories of the running beams
    int* pool_seq;                      // [2][B*nb][seq_ld] finished hypotheses
    int* anc;                           // [2][B*nb][seq_ld] K/V-cache ancestry (row that computed position j)
    float* running_score;               // [B*nb] summed log-probs
    float* pool_score;                  // [B*nb] length-penalised scores (-1e9 = empty)
    int* pool_len;                      // [B*nb] generated tokens incl. EOS
    unsigned char* pool_fin;            // [B*nb]
    unsigned char* win_done;            // [B] stop heuristic satisfied
    int* last_buf;                      // [B] buffer parity holding the window's latest pool
    float* len_pow;                     // [1024] g ** length_penalty
    int* next_token;                    // [B*nb] decoder input of the next position
    unsigned char* finished;            // [B*nb] row-skip flags of the decode kernels
    const int* step_ptr;
    int* n_active;                      // windows still searching
};
size_t beam_state_bytes(int rows, int seq_ld);
void beam_state_carve(BeamState* st, char* p, int rows, int seq_ld);
int beam_set_length_penalty(const BeamState& st, float length_penalty, int max_length, cudaStream_t stream);
int beam_init(const BeamState& st, const int* prompt_dev, cudaStream_t stream);
int beam_step(const BeamState& st, cudaStream_t stream);
int beam_output(const BeamState& st, int* tokens_out, float* scores_out, int max_new, cudaStream_t stream);
int step_increment(int* step_ptr, cudaStream_t stream);

}  // namespace wsb
