// K5b -- beam search bookkeeping on the device (the reference's default decode mode, num_beams=4:
// reference model.py:409, 614, 662 -> HF generate -> transformers/generation/utils.py `_beam_search`,
// `_get_top_k_continuations`, `_get_running_beams_for_next_iteration`, `_update_finished_beams`,
// `_check_early_stop_heuristic`).
//
// A window owns `nb` consecutive decode rows (running beams) and a pool of `nb` finished hypotheses.  Per
// position, after the tied output projection has written raw fp32 logits [rows][V]:
//   beam_row_kernel     one CTA per row: log-softmax over the full vocabulary (fp32, before the suppression
//                       masks, as HF applies its processors to log-probs), + the beam's running score, top
//                       K = 2*nb continuations of the row (value desc, token id asc);
//   beam_window_kernel  one CTA per window: merge the rows' candidates into the window's top K, split EOS /
//                       max_length hits from continuing beams, update the finished pool with the
//                       length-penalised score, evaluate the stop heuristic, and re-thread the token
//                       histories and the K/V-cache ancestry of the surviving beams.
// The self-attention K/V cache is never reordered: row r's keys at position j live in the row that computed
// them, `anc[r][j]`, and decode_attention_kernel follows that table (HF `_reorder_cache` by indirection).
// All state is double-buffered by the parity of the decoder position, so one CUDA graph serves every step.
#include "common.cuh"
#include "wsb_internal.h"
#include "decode.h"

#include <cmath>
#include <vector>

namespace wsb {

constexpr int kBeamThreads = 256;
constexpr float kNegBig = -1.0e9f;

// ---------------------------------------------------------------------------------------------- row kernel
__device__ __forceinline__ bool cand_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __launch_bounds__(kBeamThreads) beam_row_kernel(const BeamState st) {
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_wait();
    pdl_launch_dependents();
    if (st.finished[row]) return;
    const int pos = *st.step_ptr;
    const float* lg = st.logits + static_cast<long long>(row) * st.ldv;
    const bool first = (pos == st.prompt_len - 1) && st.begin_suppress != nullptr;
    __shared__ float s_a[kBeamThreads / 32], s_b[kBeamThreads / 32];
    __shared__ int s_i[kBeamThreads / 32];
    __shared__ float s_bcast[2];
    __shared__ int s_bi;

    // pass 1: max and sum of exp over the raw logits (torch log_softmax: x - max - log(sum(exp(x - max))))
    float mx = -INFINITY;
    for (int v = tid; v < st.V; v += kBeamThreads) mx = fmaxf(mx, lg[v]);
    mx = warp_max(mx);
    if (lane == 0) s_a[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float m = s_a[0];
        for (int w = 1; w < kBeamThreads / 32; ++w) m = fmaxf(m, s_a[w]);
        s_bcast[0] = m;
    }
    __syncthreads();
    mx = s_bcast[0];
    float sum = 0.0f;
    for (int v = tid; v < st.V; v += kBeamThreads) sum += expf(lg[v] - mx);
    sum = warp_sum(sum);
    if (lane == 0) s_b[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (int w = 0; w < kBeamThreads / 32; ++w) s += s_b[w];
        s_bcast[1] = logf(s);
    }
    __syncthreads();
    const float lse = s_bcast[1];
    const float base = st.running_score[row];

    // pass 2: per-thread sorted top-8 of (log-prob + mask + running score)
    float tv[8];
    int ti[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        tv[k] = -INFINITY;
        ti[k] = 0x7fffffff;
    }
    for (int v = tid; v < st.V; v += kBeamThreads) {
        float lp = (lg[v] - mx) - lse;
        lp += __ldg(st.suppress + v);
        if (first) lp += __ldg(st.begin_suppress + v);
        const float a = lp + base;
        if (cand_better(a, v, tv[7], ti[7])) {
            tv[7] = a;
            ti[7] = v;
#pragma unroll
            for (int k = 7; k > 0; --k) {
                if (cand_better(tv[k], ti[k], tv[k - 1], ti[k - 1])) {
                    const float fv = tv[k];
                    tv[k] = tv[k - 1];
                    tv[k - 1] = fv;
                    const int iv = ti[k];
                    ti[k] = ti[k - 1];
                    ti[k - 1] = iv;
                }
            }
        }
    }
    // K rounds of block arg-max over the thread heads (each list is sorted: its head is element `hd`)
    int hd = 0;
    for (int k = 0; k < st.K; ++k) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q == hd) {
                bv = tv[q];
                bi = ti[q];
            }
        if (hd >= 8) {
            bv = -INFINITY;
            bi = 0x7fffffff;
        }
        float wv = bv;
        int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (cand_better(ov, oi, wv, wi)) {
                wv = ov;
                wi = oi;
            }
        }
        if (lane == 0) {
            s_a[warp] = wv;
            s_i[warp] = wi;
        }
        __syncthreads();
        if (tid == 0) {
            float gv = s_a[0];
            int gi = s_i[0];
            for (int w = 1; w < kBeamThreads / 32; ++w)
                if (cand_better(s_a[w], s_i[w], gv, gi)) {
                    gv = s_a[w];
                    gi = s_i[w];
                }
            st.cand_val[row * 8 + k] = gv;
            st.cand_idx[row * 8 + k] = gi;
            s_bi = gi;
        }
        __syncthreads();
        if (bi == s_bi && bi != 0x7fffffff) ++hd;          // the owner pops its head
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------- window kernel
__global__ void __launch_bounds__(128) beam_window_kernel(const BeamState st) {
    const int w = blockIdx.x, tid = threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    if (st.win_done[w]) return;
    const int nb = st.nb, K = st.K;
    const int pos = *st.step_ptr;                 // decoder position of the token just consumed
    const int cur_len = pos + 1;                  // tokens in every running sequence
    const int gen_len = cur_len + 1 - st.prompt_len;
    const int ob = pos & 1, nbuf = ob ^ 1;
    const long long R = static_cast<long long>(st.B) * nb;
    const int row0 = w * nb;

    __shared__ float c_val[8];                    // window top-K (descending)
    __shared__ int c_beam[8], c_tok[8];
    __shared__ int nxt[4];                        // continuation index of every next running beam
    __shared__ int pool_src[4];                   // >=0: old pool slot, <0: -(k+1) = continuation k
    if (tid == 0) {
        // merge nb x K row candidates -> top K by (value desc, flat index beam*V+tok asc)
        float v[32];
        int f[32];
        int n = 0;
        for (int j = 0; j < nb; ++j)
            for (int k = 0; k < K; ++k) {
                v[n] = st.cand_val[(row0 + j) * 8 + k];
                f[n] = j * st.V + st.cand_idx[(row0 + j) * 8 + k];
                ++n;
            }
        bool used[32];
        for (int i = 0; i < n; ++i) used[i] = false;
        bool hit[8];
        float run_lp[8];
        for (int k = 0; k < K; ++k) {
            int best = -1;
            for (int i = 0; i < n; ++i)
                if (!used[i] && (best < 0 || v[i] > v[best] || (v[i] == v[best] && f[i] < f[best]))) best = i;
            used[best] = true;
            c_val[k] = v[best];
            c_beam[k] = f[best] / st.V;
            c_tok[k] = f[best] % st.V;
            hit[k] = (c_tok[k] == st.eos_id) || (cur_len + 1 >= st.max_length);
            run_lp[k] = c_val[k] + (hit[k] ? kNegBig : 0.0f);
        }
        // next running beams: top nb of run_lp, earlier continuation first among equals
        bool taken[8];
        for (int k = 0; k < K; ++k) taken[k] = false;
        for (int i = 0; i < nb; ++i) {
            int best = -1;
            for (int k = 0; k < K; ++k)
                if (!taken[k] && (best < 0 || run_lp[k] > run_lp[best])) best = k;
            taken[best] = true;
            nxt[i] = best;
        }
        // finished pool: old slots + the hits among the top nb continuations, keep the best nb
        const float denom = st.len_pow[gen_len];
        float m_score[12];
        int m_src[12], m_len[12];
        bool m_fin[12];
        for (int i = 0; i < nb; ++i) {
            m_score[i] = st.pool_score[row0 + i];
            m_src[i] = i;
            m_len[i] = st.pool_len[row0 + i];
            m_fin[i] = st.pool_fin[row0 + i] != 0;
        }
        for (int k = 0; k < K; ++k) {
            const bool just = hit[k] && k < nb;
            float sc = c_val[k] / denom;
            sc += just ? 0.0f : kNegBig;
            m_score[nb + k] = sc;
            m_src[nb + k] = -(k + 1);
            m_len[nb + k] = gen_len;
            m_fin[nb + k] = just;
        }
        bool mt[12];
        for (int i = 0; i < nb + K; ++i) mt[i] = false;
        float new_score[4];
        int new_len[4];
        bool new_fin[4];
        for (int s = 0; s < nb; ++s) {
            int best = -1;
            for (int i = 0; i < nb + K; ++i)
                if (!mt[i] && (best < 0 || m_score[i] > m_score[best])) best = i;
            mt[best] = true;
            pool_src[s] = m_src[best];
            new_score[s] = m_score[best];
            new_len[s] = m_len[best];
            new_fin[s] = m_fin[best];
        }
        float worst = new_score[0];
        for (int s = 0; s < nb; ++s) {
            st.pool_score[row0 + s] = new_score[s];
            st.pool_len[row0 + s] = new_len[s];
            st.pool_fin[row0 + s] = new_fin[s] ? 1 : 0;
            worst = fminf(worst, new_score[s]);
        }
        // running state + stop heuristic (early_stopping=False)
        for (int i = 0; i < nb; ++i) {
            st.running_score[row0 + i] = run_lp[nxt[i]];
            st.next_token[row0 + i] = c_tok[nxt[i]];
        }
        const float best_possible = run_lp[nxt[0]] / denom;
        bool unsat = false;
        for (int s = 0; s < nb; ++s) unsat = unsat || (best_possible > (new_fin[s] ? worst : kNegBig));
        st.last_buf[w] = nbuf;
        if (!unsat) {
            st.win_done[w] = 1;
            for (int i = 0; i < nb; ++i) st.finished[row0 + i] = 1;
            atomicSub(st.n_active, 1);
        }
    }
    __syncthreads();
    // re-thread histories: new beam i continues old beam c_beam[nxt[i]]
    const long long seq_o = static_cast<long long>(ob) * R * st.seq_ld, seq_n = static_cast<long long>(nbuf) * R * st.seq_ld;
    for (int i = 0; i < nb; ++i) {
        const int k = nxt[i];
        const int parent = row0 + c_beam[k];
        const int* so = st.run_seq + seq_o + static_cast<long long>(parent) * st.seq_ld;
        int* sn = st.run_seq + seq_n + static_cast<long long>(row0 + i) * st.seq_ld;
        for (int j = tid; j < cur_len; j += blockDim.x) sn[j] = so[j];
        const int* ao = st.anc + seq_o + static_cast<long long>(parent) * st.seq_ld;
        int* an = st.anc + seq_n + static_cast<long long>(row0 + i) * st.seq_ld;
        for (int j = tid; j < pos; j += blockDim.x) an[j] = ao[j];
        if (tid == 0) {
            sn[cur_len] = c_tok[k];
            an[pos] = parent;
        }
    }
    for (int s = 0; s < nb; ++s) {
        int* pn = st.pool_seq + seq_n + static_cast<long long>(row0 + s) * st.seq_ld;
        const int src = pool_src[s];
        if (src >= 0) {
            const int* po = st.pool_seq + seq_o + static_cast<long long>(row0 + src) * st.seq_ld;
            const int n = st.prompt_len + st.pool_len[row0 + s];
            for (int j = tid; j < n; j += blockDim.x) pn[j] = po[j];
        } else {
            const int k = -src - 1;
            const int* so = st.run_seq + seq_o + static_cast<long long>(row0 + c_beam[k]) * st.seq_ld;
            for (int j = tid; j < cur_len; j += blockDim.x) pn[j] = so[j];
            if (tid == 0) pn[cur_len] = c_tok[k];
        }
    }
}

__global__ void beam_init_kernel(const BeamState st, const int* __restrict__ prompt) {
    const long long R = static_cast<long long>(st.B) * st.nb;
    const long long n = 2 * R * st.seq_ld;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(i % st.seq_ld);
        const int r = static_cast<int>((i / st.seq_ld) % R);
        const int tok = j < st.prompt_len ? prompt[j] : st.pad_id;
        st.run_seq[i] = tok;
        st.pool_seq[i] = tok;
        st.anc[i] = r;
    }
    for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < R;
         r += static_cast<long long>(gridDim.x) * blockDim.x) {
        st.running_score[r] = (r % st.nb) == 0 ? 0.0f : kNegBig;
        st.pool_score[r] = kNegBig;
        st.pool_len[r] = 0;
        st.pool_fin[r] = 0;
        st.finished[r] = 0;
        st.next_token[r] = prompt[0];
        if (r < st.B) {
            st.win_done[r] = 0;
            st.last_buf[r] = 0;
        }
    }
}

// best hypothesis of every window, prompt stripped, padded: tokens_out [B][max_new]; scores_out [B] (nullable)
__global__ void beam_output_kernel(const BeamState st, int* __restrict__ tokens_out, float* __restrict__ scores_out,
                                   int max_new) {
    const int w = blockIdx.x;
    const long long R = static_cast<long long>(st.B) * st.nb;
    const int* seq = st.pool_seq + (static_cast<long long>(st.last_buf[w]) * R + static_cast<long long>(w) * st.nb) * st.seq_ld;
    const int len = st.pool_fin[w * st.nb] ? st.pool_len[w * st.nb] : 0;
    for (int i = threadIdx.x; i < max_new; i += blockDim.x)
        tokens_out[static_cast<long long>(w) * max_new + i] = i < len ? seq[st.prompt_len + i] : st.pad_id;
    if (scores_out && threadIdx.x == 0) scores_out[w] = st.pool_score[w * st.nb];
}

size_t beam_state_bytes(int rows, int seq_ld) {
    const size_t R = rows;
    auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
    return al(R * 8 * 4) * 2 + al(2 * R * seq_ld * 4) * 3 + al(R * 4) * 3 + al(R) * 2 + al(R * 4) + al(R) + al(1024 * 4);
}

void beam_state_carve(BeamState* st, char* p, int rows, int seq_ld) {
    const size_t R = rows;
    auto take = [&](size_t bytes) {
        char* r = p;
        p += (bytes + 255) & ~size_t(255);
        return r;
    };
    st->cand_val = reinterpret_cast<float*>(take(R * 8 * 4));
    st->cand_idx = reinterpret_cast<int*>(take(R * 8 * 4));
    st->run_seq = reinterpret_cast<int*>(take(2 * R * seq_ld * 4));
    st->pool_seq = reinterpret_cast<int*>(take(2 * R * seq_ld * 4));
    st->anc = reinterpret_cast<int*>(take(2 * R * seq_ld * 4));
    st->running_score = reinterpret_cast<float*>(take(R * 4));
    st->pool_score = reinterpret_cast<float*>(take(R * 4));
    st->pool_len = reinterpret_cast<int*>(take(R * 4));
    st->pool_fin = reinterpret_cast<unsigned char*>(take(R));
    st->win_done = reinterpret_cast<unsigned char*>(take(R));
    st->last_buf = reinterpret_cast<int*>(take(R * 4));
    st->len_pow = reinterpret_cast<float*>(take(1024 * 4));
    st->seq_ld = seq_ld;
}

int beam_set_length_penalty(const BeamState& st, float length_penalty, int max_length, cudaStream_t stream) {
    WSB_REQUIRE(max_length < 1024, "max_length < 1024");
    std::vector<float> tab(1024, 1.0f);
    for (int g = 1; g < 1024; ++g) tab[g] = static_cast<float>(std::pow(static_cast<double>(g), static_cast<double>(length_penalty)));
    WSB_CHECK_CUDA(cudaMemcpyAsync(st.len_pow, tab.data(), sizeof(float) * 1024, cudaMemcpyHostToDevice, stream));
    WSB_CHECK_CUDA(cudaStreamSynchronize(stream));         // tab is a stack temporary
    return 0;
}

int beam_init(const BeamState& st, const int* prompt_dev, cudaStream_t stream) {
    beam_init_kernel<<<148, 256, 0, stream>>>(st, prompt_dev);
    WSB_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int beam_step(const BeamState& st, cudaStream_t stream) {
    WSB_REQUIRE(st.nb >= 1 && st.nb <= 4 && st.K == 2 * st.nb, "num_beams in [1,4]");
    WSB_CHECK_CUDA(launch_kernel(beam_row_kernel, dim3(st.B * st.nb), dim3(kBeamThreads), 0, stream, st));
    WSB_CHECK_CUDA(launch_kernel(beam_window_kernel, dim3(st.B), dim3(128), 0, stream, st));
    count_launch(2);
    return 0;
}

int beam_output(const BeamState& st, int* tokens_out, float* scores_out, int max_new, cudaStream_t stream) {
    beam_output_kernel<<<st.B, 128, 0, stream>>>(st, tokens_out, scores_out, max_new);
    WSB_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace wsb
