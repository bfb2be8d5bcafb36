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
int decode_self_attention(const __nv_bfloat16* qkv, const SplitkInput* part, int d, __nv_bfloat16* k_cache,
                          __nv_bfloat16* v_cache, int t_max, const int* step_ptr, int pos_offset,
                          const unsigned char* finished, __nv_bfloat16* out, int B, int n_heads, cudaStream_t stream);
int decode_cross_attention(const __nv_bfloat16* q, const SplitkInput* part, int d, const __nv_bfloat16* cross_kv, int layer,
                           int n_layers, int T, const unsigned char* finished, __nv_bfloat16* out, int B, int n_heads,
                           cudaStream_t stream);
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

}  // namespace wsb
