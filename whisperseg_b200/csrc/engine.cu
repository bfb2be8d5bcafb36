// Engine: owns the per-model workspace and sequences the kernels of the hot path.
//   wsb_encode   -- conv stem + encoder layers           (HF WhisperEncoder.forward, modeling_whisper.py:593-647)
//   wsb_generate -- cross-K/V projection + greedy KV-cache decode loop, optionally replayed from a
//                   CUDA graph                            (HF generate, reference model.py:655-666)
// plus the extern "C" surface declared in include/wsb.h.
#include "common.cuh"
#include "wsb_internal.h"
#include "decode.h"
#include "../../include/wsb.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <tuple>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace wsb {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
thread_local bool g_use_pdl = false;
thread_local int g_sm_reserve = 0;

// ---------------------------------------------------------------------------- event profiling
// Optional CUDA-event bracketing of kernel classes on the launching stream (bench.py's roofline
// numbers).  Only eager launches are bracketed (never inside a graph capture).
enum ProfCat : int {
    PROF_CONV1 = 0, PROF_ENC_GEMM, PROF_ENC_ATTN, PROF_ENC_LN, PROF_CROSSKV_GEMM, PROF_DEC_GEMM, PROF_DEC_LOGITS,
    PROF_DEC_SELF_ATTN, PROF_DEC_CROSS_ATTN, PROF_DEC_LN, PROF_DEC_MISC, PROF_DEC_GRAPH, PROF_DEC_COMPACT, PROF_NCAT
};
struct ProfRec {
    cudaEvent_t a, b;
    int cat;
    double work;
};
struct Profiler {
    bool enabled = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    double ms[PROF_NCAT] = {0};
    double work[PROF_NCAT] = {0};
    long long count[PROF_NCAT] = {0};
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void collect() {
        for (auto& r : recs) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
                ms[r.cat] += t;
                work[r.cat] += r.work;
                count[r.cat] += 1;
            }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
};
static Profiler g_prof;
static bool g_capturing = false;
static std::mutex g_mega_mutex[64];
struct ProfScope {
    cudaStream_t s;
    bool on;
    ProfRec r;
    ProfScope(int cat, double work, cudaStream_t stream) : s(stream), on(g_prof.enabled && !g_capturing) {
        if (on) {
            r.a = g_prof.get();
            r.b = g_prof.get();
            r.cat = cat;
            r.work = work;
            cudaEventRecord(r.a, s);
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(r.b, s);
            g_prof.recs.push_back(r);
        }
    }
};

struct EncLayer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    const __nv_bfloat16 *qkv_w, *o_w, *fc1_w, *fc2_w;
    const float *qkv_b, *o_b, *fc1_b, *fc2_b;
};
struct DecLayer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
    const __nv_bfloat16 *sqkv_w, *so_w, *cq_w, *co_w, *fc1_w, *fc2_w;
    const float *sqkv_b, *so_b, *cq_b, *co_b, *fc1_b, *fc2_b;
    // optional (all or none): the three LayerNorm-consuming projections with the LayerNorm affine folded in --
    // *_wf = bf16(W o gamma), *_c1[n] = sum_k wf[n][k], *_c2 = b + W beta (weights.py: fold_layernorm)
    const __nv_bfloat16 *sqkv_wf = nullptr, *cq_wf = nullptr, *fc1_wf = nullptr;
    const float *sqkv_c1 = nullptr, *sqkv_c2 = nullptr, *cq_c1 = nullptr, *cq_c2 = nullptr, *fc1_c1 = nullptr, *fc1_c2 = nullptr;
};

struct Model {
    wsb_model_config cfg;
    int T;                                // encoder positions = n_cols / 2
    // encoder weights
    const float *conv1_wt, *conv1_b, *conv2_b, *enc_pos, *enc_ln_g, *enc_ln_b;
    const __nv_bfloat16* conv2_w;
    const __nv_bfloat16* conv1_wg = nullptr;   // optional: conv1 as a GEMM (weights.py: enc.conv1.wg); WSB_CONV1_FP32=1 keeps the CUDA-core kernel
    __nv_bfloat16* feat_tm = nullptr;          // [B][n_cols + 3][160] time-major bf16 features (hi | lo)
    std::vector<EncLayer> enc;
    // decoder weights
    const __nv_bfloat16 *dec_emb, *crosskv_w;
    const float *dec_pos, *crosskv_b, *dec_ln_g, *dec_ln_b, *suppress, *begin_suppress;
    std::vector<DecLayer> dec;
    // workspace
    char* ws = nullptr;
    size_t ws_bytes = 0;
    __nv_bfloat16 *h1p, *xn, *qkv, *att, *ff, *enc_out, *cross_kv, *k_cache, *v_cache, *dxn, *dqkv, *datt, *dq, *dff;
    float *x, *dx, *am_val, *dpart, *gv_stats;
    size_t dpart_floats = 0;
    int *am_idx, *tokens, *next_token, *step, *n_active, *prompt_dev;
    unsigned char* finished;
    int am_tiles = 0, logits_bn = 0;
    int* logit_tiles = nullptr;          // n-tiles of the output projection that contain a non-suppressed token
    int last_batch = 0;
    // compact decode state (see decode.cu: batch compaction): alternate K/V buffers for kCompactRows rows
    __nv_bfloat16 *cross_kv_alt, *k_cache_alt, *v_cache_alt;
    int *next_token_alt, *row_map_main, *row_map_alt, *active_idx;
    unsigned char* finished_alt;
    // CUDA graphs of one steady-state decode step, keyed by everything that is baked into the launches
    struct GraphEntry {
        cudaGraphExec_t exec;
        int kernels;
    };
    std::map<std::tuple<int, int, int, int, int, int, int>, GraphEntry> graphs;
    bool use_pdl = false;
    bool use_gemv = true;               // few decode rows: fused LN + mma.sync linear layers (gemv.cu)
    bool use_cluster = false;           // 65..256 decode rows: cluster split-K linear layers with folded LayerNorm (skinny.cu);
                                        // opt-in (WSB_CLUSTER=1): parity-green, but measured 0-4 % slower than the split-K pair
    bool use_fold = true;               // ... with the LayerNorm folded into the projection when the folded tensors exist
    int wide_direct_bn = 128;           // > 64 rows: qkv / cross-q / fc1 as single full-K launches with this block_n (one fp32 plane, or
                                        // bias + GELU in the GEMM epilogue) instead of split-K planes + a second phase; WSB_WIDE_DIRECT=0: off
    int wide_cq_bn = 32;                // ... block_n of the cross-q projection on that path (WSB_WIDE_CQ_BN)
    int wide_direct_resid = 0;          // > 64 rows: self-out / cross-out as one full-K launch (block_n) + LayerNorm kernel (WSB_WIDE_DIRECT_RESID)
    bool use_mega = false;              // <= 64 rows: one persistent kernel per decoder position (mega.cu); opt-in (WSB_MEGA=1):
                                        // bit-identical tokens, but measured 3x SLOWER than the launch-per-layer path
    void* mega_layers = nullptr;        // device table of the folded linear layers (null: folded tensors missing / unsupported width)
    unsigned int* mega_sync = nullptr;  // device: grid-barrier arrivals, exits, watchdog flag
    unsigned long long* mega_trace = nullptr;   // diagnostics (wsb_mega_trace): per-barrier timestamps of the last position
    bool fold_guard = true;             // fall back to the exact LayerNorm when a row's common mode dominates (WSB_FOLD_GUARD=0: off)
    bool fold_disabled = false;         // sticky: the guard fired once for this model
    int gemv_rows = 64;                 // ... used up to this many rows (WSB_GEMV_ROWS, <= 64)
    std::vector<int> ladder{64, 32, 16};   // compaction levels, descending, all <= kCompactRows (override: WSB_LADDER=64,16)
    int* pinned_active = nullptr;
    // beam search workspace (allocated on first use): raw logits [max_batch][ldv] + BeamState arrays
    char* beam_ws = nullptr;
    float* beam_logits = nullptr;
    long long beam_ldv = 0;
    BeamState beam;
    std::map<std::tuple<int, int, int, int, int, int, int>, GraphEntry> beam_graphs;
};

constexpr size_t kPrefillRows = 4;        // prompt positions decoded in one pass (prompt_len <= 4; the reference's prompt has 3 tokens)
constexpr size_t kMaxCachedGraphs = 48;   // instantiated decode-step graphs kept per model (each distinct batch size x ladder level is one)
constexpr int kCompactRows = 64;        // first compaction level; the second level (16 rows) reuses the main buffers

// the per-row decode state a step works on (main buffers, or a compacted copy)
struct DecState {
    int B;
    int buffer_id;                      // 0 = main buffers, 1 = alternate buffers
    __nv_bfloat16 *k_cache, *v_cache, *cross_kv;
    int* next_token;
    unsigned char* finished;
    const int* row_map;                 // slot -> window index of this generate() call (null = identity)
    int kv_div = 1;                     // beam search: rows per window (shared cross-attention K/V block)
    const int* anc = nullptr;           // beam search: K/V-cache ancestry tables
    int anc_ld = 0;
};

template <typename T>
static T* carve(char*& p, size_t count) {
    T* r = reinterpret_cast<T*>(p);
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    p += bytes;
    return r;
}

static int model_layout(Model* m, bool assign) {
    const wsb_model_config& c = m->cfg;
    const size_t B = c.max_batch, d = c.d_model, F = c.ffn_dim, L = c.n_layers, H = c.n_heads, T = m->T;
    const size_t rows = B * T;
    char* p = assign ? m->ws : nullptr;
    char* p0 = p;
    m->h1p = carve<__nv_bfloat16>(p, B * (c.n_cols + 1) * d);
    m->feat_tm = carve<__nv_bfloat16>(p, B * (c.n_cols + 3) * 160 + 256);
    m->x = carve<float>(p, rows * d);
    m->xn = carve<__nv_bfloat16>(p, rows * d);
    m->qkv = carve<__nv_bfloat16>(p, rows * 3 * d);
    m->att = carve<__nv_bfloat16>(p, rows * d);
    m->ff = carve<__nv_bfloat16>(p, rows * F);
    m->enc_out = carve<__nv_bfloat16>(p, rows * d);
    m->cross_kv = carve<__nv_bfloat16>(p, B * L * 2 * H * T * 64);
    m->k_cache = carve<__nv_bfloat16>(p, L * B * H * c.max_target_positions * 64);
    m->v_cache = carve<__nv_bfloat16>(p, L * B * H * c.max_target_positions * 64);
    // (the row-indexed decode scratch holds kPrefillRows x B rows: the prompt positions run as one pass of P x B virtual rows)
    m->dx = carve<float>(p, kPrefillRows * B * d);
    m->dxn = carve<__nv_bfloat16>(p, kPrefillRows * B * d);
    m->dqkv = carve<__nv_bfloat16>(p, B * 3 * d);
    m->datt = carve<__nv_bfloat16>(p, kPrefillRows * B * d);
    m->dq = carve<__nv_bfloat16>(p, B * d);
    m->dff = carve<__nv_bfloat16>(p, kPrefillRows * B * F);
    m->dpart_floats = 16 * B * std::max<size_t>(F, 3 * d);
    m->dpart = carve<float>(p, m->dpart_floats);
    m->gv_stats = carve<float>(p, std::max<size_t>(256 * 64 * 2, static_cast<size_t>(c.d_model / 128 + 1) * B * 2));
    m->am_val = carve<float>(p, B * m->am_tiles);
    m->am_idx = carve<int>(p, B * m->am_tiles);
    m->tokens = carve<int>(p, B * c.max_target_positions);
    m->next_token = carve<int>(p, B);
    m->step = carve<int>(p, 4);
    m->n_active = carve<int>(p, 4);
    m->prompt_dev = carve<int>(p, 16);
    m->finished = carve<unsigned char>(p, B);
    {
        const size_t Ba = std::min<size_t>(B, kCompactRows);
        m->cross_kv_alt = carve<__nv_bfloat16>(p, Ba * L * 2 * H * T * 64);
        m->k_cache_alt = carve<__nv_bfloat16>(p, L * Ba * H * c.max_target_positions * 64);
        m->v_cache_alt = carve<__nv_bfloat16>(p, L * Ba * H * c.max_target_positions * 64);
        m->next_token_alt = carve<int>(p, B);
        m->row_map_main = carve<int>(p, B);
        m->row_map_alt = carve<int>(p, B);
        m->active_idx = carve<int>(p, B);
        m->finished_alt = carve<unsigned char>(p, B);
    }
    m->ws_bytes = static_cast<size_t>(p - p0);
    return 0;
}

static int lookup(const std::unordered_map<std::string, const void*>& tab, const std::string& name, const void** out) {
    auto it = tab.find(name);
    if (it == tab.end() || it->second == nullptr) {
        set_last_error("wsb_model_create: missing tensor '" + name + "'");
        return 5;
    }
    *out = it->second;
    return 0;
}
#define WSB_GET(field, name)                                                       \
    do {                                                                           \
        const void* _p = nullptr;                                                  \
        int _rc = lookup(tab, name, &_p);                                          \
        if (_rc) return _rc;                                                       \
        field = reinterpret_cast<decltype(field)>(_p);                             \
    } while (0)

static int model_create(const wsb_model_config* cfg, const char* const* names, const void* const* ptrs, int n,
                        Model** out) {
    WSB_REQUIRE(cfg->d_model == cfg->n_heads * 64, "head_dim must be 64");
    WSB_REQUIRE(cfg->d_model % 64 == 0 && cfg->ffn_dim % 64 == 0, "d_model / ffn_dim must be multiples of 64");
    WSB_REQUIRE(cfg->n_cols % 2 == 0 && cfg->n_cols / 2 <= 512, "n_cols/2 encoder positions must be <= 512");
    WSB_REQUIRE(cfg->n_mels == 80, "80 mel bins");
    WSB_REQUIRE(cfg->max_batch >= 1, "max_batch >= 1");
    WSB_REQUIRE(cfg->max_target_positions <= 512, "max_target_positions <= 512");
    int dev = 0, major = 0;
    WSB_CHECK_CUDA(cudaGetDevice(&dev));
    WSB_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    WSB_REQUIRE(major == 10, "libwsb needs an sm_100 (Blackwell B200) device; there is no fallback path");
    std::unordered_map<std::string, const void*> tab;
    for (int i = 0; i < n; ++i) tab[names[i]] = ptrs[i];
    Model* m = new Model();
    m->cfg = *cfg;
    m->T = cfg->n_cols / 2;
    WSB_GET(m->conv1_wt, "enc.conv1.wt");
    WSB_GET(m->conv1_b, "enc.conv1.b");
    WSB_GET(m->conv2_w, "enc.conv2.w");
    if (tab.count("enc.conv1.wg")) WSB_GET(m->conv1_wg, "enc.conv1.wg");
    WSB_GET(m->conv2_b, "enc.conv2.b");
    WSB_GET(m->enc_pos, "enc.pos");
    WSB_GET(m->enc_ln_g, "enc.ln.g");
    WSB_GET(m->enc_ln_b, "enc.ln.b");
    m->enc.resize(cfg->n_layers);
    for (int l = 0; l < cfg->n_layers; ++l) {
        const std::string p = "enc." + std::to_string(l) + ".";
        EncLayer& e = m->enc[l];
        WSB_GET(e.ln1_g, p + "ln1.g"); WSB_GET(e.ln1_b, p + "ln1.b");
        WSB_GET(e.qkv_w, p + "qkv.w"); WSB_GET(e.qkv_b, p + "qkv.b");
        WSB_GET(e.o_w, p + "o.w"); WSB_GET(e.o_b, p + "o.b");
        WSB_GET(e.ln2_g, p + "ln2.g"); WSB_GET(e.ln2_b, p + "ln2.b");
        WSB_GET(e.fc1_w, p + "fc1.w"); WSB_GET(e.fc1_b, p + "fc1.b");
        WSB_GET(e.fc2_w, p + "fc2.w"); WSB_GET(e.fc2_b, p + "fc2.b");
    }
    WSB_GET(m->dec_emb, "dec.emb");
    WSB_GET(m->dec_pos, "dec.pos");
    WSB_GET(m->crosskv_w, "dec.crosskv.w");
    WSB_GET(m->crosskv_b, "dec.crosskv.b");
    WSB_GET(m->dec_ln_g, "dec.ln.g");
    WSB_GET(m->dec_ln_b, "dec.ln.b");
    WSB_GET(m->suppress, "dec.suppress");
    WSB_GET(m->begin_suppress, "dec.begin_suppress");
    m->dec.resize(cfg->n_layers);
    for (int l = 0; l < cfg->n_layers; ++l) {
        const std::string p = "dec." + std::to_string(l) + ".";
        DecLayer& e = m->dec[l];
        WSB_GET(e.ln1_g, p + "ln1.g"); WSB_GET(e.ln1_b, p + "ln1.b");
        WSB_GET(e.sqkv_w, p + "sqkv.w"); WSB_GET(e.sqkv_b, p + "sqkv.b");
        WSB_GET(e.so_w, p + "so.w"); WSB_GET(e.so_b, p + "so.b");
        WSB_GET(e.ln2_g, p + "ln2.g"); WSB_GET(e.ln2_b, p + "ln2.b");
        WSB_GET(e.cq_w, p + "cq.w"); WSB_GET(e.cq_b, p + "cq.b");
        WSB_GET(e.co_w, p + "co.w"); WSB_GET(e.co_b, p + "co.b");
        WSB_GET(e.ln3_g, p + "ln3.g"); WSB_GET(e.ln3_b, p + "ln3.b");
        WSB_GET(e.fc1_w, p + "fc1.w"); WSB_GET(e.fc1_b, p + "fc1.b");
        WSB_GET(e.fc2_w, p + "fc2.w"); WSB_GET(e.fc2_b, p + "fc2.b");
        if (tab.count(p + "sqkv.wf")) {
            WSB_GET(e.sqkv_wf, p + "sqkv.wf"); WSB_GET(e.sqkv_c1, p + "sqkv.c1"); WSB_GET(e.sqkv_c2, p + "sqkv.c2");
            WSB_GET(e.cq_wf, p + "cq.wf"); WSB_GET(e.cq_c1, p + "cq.c1"); WSB_GET(e.cq_c2, p + "cq.c2");
            WSB_GET(e.fc1_wf, p + "fc1.wf"); WSB_GET(e.fc1_c1, p + "fc1.c1"); WSB_GET(e.fc1_c2, p + "fc1.c2");
        }
    }
    m->logits_bn = gemm_pick_block_n(cfg->max_batch, cfg->vocab_size);
    {   // vocabulary tiles in which every token is suppressed (additive -inf mask) can never win the arg-max:
        // only the others are computed
        std::vector<float> mask(cfg->vocab_size);
        WSB_CHECK_CUDA(cudaMemcpy(mask.data(), m->suppress, sizeof(float) * cfg->vocab_size, cudaMemcpyDeviceToHost));
        std::vector<int> tiles;
        const int nt = gemm_n_tiles(cfg->vocab_size, m->logits_bn);
        for (int t = 0; t < nt; ++t) {
            bool any = false;
            for (int v = t * m->logits_bn; v < std::min(cfg->vocab_size, (t + 1) * m->logits_bn) && !any; ++v)
                any = !(mask[v] < -1e30f);
            if (any) tiles.push_back(t);
        }
        if (tiles.empty()) tiles.push_back(0);
        m->am_tiles = static_cast<int>(tiles.size());
        WSB_CHECK_CUDA(cudaMalloc(&m->logit_tiles, sizeof(int) * tiles.size()));
        WSB_CHECK_CUDA(cudaMemcpy(m->logit_tiles, tiles.data(), sizeof(int) * tiles.size(), cudaMemcpyHostToDevice));
    }
    model_layout(m, false);
    WSB_CHECK_CUDA(cudaMalloc(&m->ws, m->ws_bytes));
    model_layout(m, true);
    WSB_CHECK_CUDA(cudaMallocHost(&m->pinned_active, sizeof(int) * 4));
    if (m->dec[0].sqkv_wf != nullptr && mega_supported(cfg->d_model, cfg->ffn_dim)) {
        std::vector<char> host(mega_layer_table_bytes(cfg->n_layers));
        for (int l = 0; l < cfg->n_layers; ++l) {
            const DecLayer& e = m->dec[l];
            const void* W[6] = {e.sqkv_wf, e.so_w, e.cq_wf, e.co_w, e.fc1_wf, e.fc2_w};
            const float* bias[6] = {e.sqkv_c2, e.so_b, e.cq_c2, e.co_b, e.fc1_c2, e.fc2_b};
            const float* c1[6] = {e.sqkv_c1, nullptr, e.cq_c1, nullptr, e.fc1_c1, nullptr};
            mega_fill_layer(host.data(), l, W, bias, c1);
        }
        WSB_CHECK_CUDA(cudaMalloc(&m->mega_layers, host.size()));
        WSB_CHECK_CUDA(cudaMemcpy(m->mega_layers, host.data(), host.size(), cudaMemcpyHostToDevice));
        WSB_CHECK_CUDA(cudaMalloc(&m->mega_sync, sizeof(unsigned int) * 64));
        WSB_CHECK_CUDA(cudaMemset(m->mega_sync, 0, sizeof(unsigned int) * 64));
    }
    *out = m;
    return 0;
}

static void model_destroy(Model* m) {
    if (!m) return;
    for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second.exec);
    for (auto& kv : m->beam_graphs) cudaGraphExecDestroy(kv.second.exec);
    cudaFree(m->beam_ws);
    cudaFree(m->mega_layers);
    cudaFree(m->mega_sync);
    cudaFree(m->mega_trace);
    cudaFree(m->ws);
    cudaFree(m->logit_tiles);
    cudaFreeHost(m->pinned_active);
    delete m;
}

#define WSB_RUN(expr)        \
    do {                     \
        int _rc = (expr);    \
        if (_rc) return _rc; \
    } while (0)

static int linear(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, int M, int N, int K, int act,
                  const float* resid, void* out, int out_mode, cudaStream_t s, int block_n = 0, int cat = PROF_ENC_GEMM) {
    ProfScope ps(cat, 2.0 * M * N * K, s);
    GemmArgs g;
    g.A = A;
    g.lda = K;
    g.W = W;
    g.M = M;
    g.N = N;
    g.K = K;
    g.bias = bias;
    g.act = act;
    g.resid = resid;
    g.ldr = N;
    g.out = out;
    g.ldc = N;
    g.out_mode = out_mode;
    g.block_n = block_n;
    return gemm_bf16(g, s);
}

static int encode(Model* m, const float* feats, int B, float* hidden_f32, cudaStream_t s) {
    const wsb_model_config& c = m->cfg;
    WSB_REQUIRE(B >= 1 && B <= c.max_batch, "batch exceeds the model's max_batch");
    const int d = c.d_model, T = m->T, rows = B * T;
    const long long h1_stride = static_cast<long long>(c.n_cols + 1) * d;
    if (m->conv1_wg != nullptr && std::getenv("WSB_CONV1_FP32") == nullptr && d % 32 == 0) {
        // conv1 (k = 3, pad 1) as an im2col-free tcgen05 GEMM: row t of batch b is the contiguous span of 3 x (80 hi + 80 lo) bf16
        // features starting at time-major row t (zero row in front), K padded to 512 with zero weights; bias + GELU in the epilogue,
        // written straight into h1p rows 1.. of every batch (row 0 = conv2's left padding)
        ProfScope ps(PROF_CONV1, 2.0 * B * c.n_cols * 240.0 * d, s);
        WSB_RUN(features_time_major_bf16(feats, m->feat_tm, B, c.n_cols, s));
        WSB_CHECK_CUDA(cudaMemset2DAsync(m->h1p, static_cast<size_t>(h1_stride) * 2, 0, static_cast<size_t>(d) * 2, B, s));
        GemmArgs g;
        g.A = m->feat_tm;
        g.lda = 160;
        g.a_rows_per_batch = c.n_cols;
        g.a_batch_stride = static_cast<long long>(c.n_cols + 3) * 160;
        g.W = m->conv1_wg;
        g.M = B * c.n_cols;
        g.N = d;
        g.K = 512;
        g.bias = m->conv1_b;
        g.act = GEMM_ACT_GELU;
        g.out = m->h1p + d;
        g.ldc = d;
        g.out_mode = GEMM_OUT_BF16;
        g.c_batch_pad = 1;
        WSB_RUN(gemm_bf16(g, s));
    } else {
        ProfScope ps(PROF_CONV1, 2.0 * B * c.n_cols * 240.0 * d, s);
        WSB_RUN(conv1_gelu(feats, m->conv1_wt, m->conv1_b, m->h1p, B, c.n_cols, d, h1_stride, s));
    }
    {   // conv2 (k=3, stride 2, pad 1) as an im2col-free GEMM: row t' of batch b is the contiguous span
        // h1p[b][2t' .. 2t'+2][:] (3d elements), i.e. a strided, overlapping view of h1p.
        GemmArgs g;
        g.A = m->h1p;
        g.lda = 2LL * d;
        g.a_rows_per_batch = T;
        g.a_batch_stride = h1_stride;
        g.W = m->conv2_w;
        g.M = rows;
        g.N = d;
        g.K = 3 * d;
        g.bias = m->conv2_b;
        g.act = GEMM_ACT_GELU;
        g.rowvec = m->enc_pos;
        g.rows_per_batch = T;
        g.out = m->x;
        g.ldc = d;
        g.out_mode = GEMM_OUT_F32;
        ProfScope ps(PROF_ENC_GEMM, 2.0 * rows * d * 3.0 * d, s);
        WSB_RUN(gemm_bf16(g, s));
    }
    for (int l = 0; l < c.n_layers; ++l) {
        const EncLayer& e = m->enc[l];
        {
            ProfScope ps(PROF_ENC_LN, 6.0 * rows * d, s);
            WSB_RUN(layernorm_f32_to_bf16(m->x, e.ln1_g, e.ln1_b, m->xn, nullptr, rows, d, s));
        }
        WSB_RUN(linear(m->xn, e.qkv_w, e.qkv_b, rows, 3 * d, d, GEMM_ACT_NONE, nullptr, m->qkv, GEMM_OUT_BF16, s));
        {
            ProfScope ps(PROF_ENC_ATTN, 4.0 * B * c.n_heads * T * T * 64.0, s);
            WSB_RUN(encoder_attention(m->qkv, m->att, B, T, c.n_heads, s));
        }
        WSB_RUN(linear(m->att, e.o_w, e.o_b, rows, d, d, GEMM_ACT_NONE, m->x, m->x, GEMM_OUT_F32, s));
        {
            ProfScope ps(PROF_ENC_LN, 6.0 * rows * d, s);
            WSB_RUN(layernorm_f32_to_bf16(m->x, e.ln2_g, e.ln2_b, m->xn, nullptr, rows, d, s));
        }
        WSB_RUN(linear(m->xn, e.fc1_w, e.fc1_b, rows, c.ffn_dim, d, GEMM_ACT_GELU, nullptr, m->ff, GEMM_OUT_BF16, s));
        WSB_RUN(linear(m->ff, e.fc2_w, e.fc2_b, rows, d, c.ffn_dim, GEMM_ACT_NONE, m->x, m->x, GEMM_OUT_F32, s));
    }
    WSB_RUN(layernorm_f32_to_bf16(m->x, m->enc_ln_g, m->enc_ln_b, m->enc_out, hidden_f32, rows, d, s));
    m->last_batch = B;
    return 0;
}

// skinny (decode) linear layer: split-K tcgen05 GEMM into fp32 partial planes, then the fused second phase
//   mode 0: out_bf16 = act(sum + bias)         mode 1: x += sum + bias; xn = LayerNorm(x) (if gamma)
static int skinny_linear(Model* m, const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, int B, int N, int K,
                         int gelu, __nv_bfloat16* out_bf16, float* x, const float* gamma, const float* beta,
                         __nv_bfloat16* xn, cudaStream_t s, const unsigned char* row_skip, SplitkInput* planes_only = nullptr,
                         int force_bn = 0) {
    int bn = 128, splits = 1;
    gemm_pick_skinny(B, N, K, &bn, &splits);
    if (force_bn > 0) {                                 // one full-K tile per CTA: a single fp32 plane, no cross-CTA sum
        bn = force_bn;
        splits = 1;
    }
    const int64_t plane = static_cast<int64_t>(B) * N;
    WSB_REQUIRE(static_cast<size_t>(plane) * splits <= m->dpart_floats, "split-K workspace too small");
    {
        ProfScope ps(PROF_DEC_GEMM, 2.0 * B * N * K, s);
        GemmArgs g;
        g.A = A;
        g.lda = K;
        g.W = W;
        g.M = B;
        g.N = N;
        g.K = K;
        g.out = m->dpart;
        g.ldc = N;
        g.out_mode = GEMM_OUT_F32;
        g.block_n = bn;
        g.splits = splits;
        g.split_stride = plane;
        g.row_skip = row_skip;
        WSB_RUN(gemm_bf16(g, s));
    }
    const int eff = gemm_effective_splits(K, splits);
    if (planes_only) {                                  // the consumer kernel performs the second phase itself
        planes_only->planes = m->dpart;
        planes_only->splits = eff;
        planes_only->split_stride = plane;
        planes_only->bias = bias;
        return 0;
    }
    ProfScope ps(PROF_DEC_LN, 4.0 * B * N * (splits + 1), s);
    if (out_bf16) return splitk_reduce_bf16(m->dpart, eff, plane, B, N, bias, gelu, out_bf16, row_skip, s);
    return splitk_reduce_resid_ln(m->dpart, eff, plane, B, N, bias, x, gamma, beta, xn, row_skip, s);
}

// what a decoder position's layer stack needs, whichever linear-layer implementation runs it
struct StepCtx {
    Model* m;
    const DecState& st;
    int B, d, F, L, H, T, tmax;
    const unsigned char* fin;            // finished-row flags (null under teacher forcing)
    long long cache_l;                   // elements of one layer's K (or V) cache
    bool with_logits;
    cudaStream_t s;
    int P = 0;                           // > 0: prompt prefill, P positions x B rows as P * B virtual rows (split-K path only)
};
#define WSB_STEP_LOCALS                                                                                      \
    Model* m = x.m;                                                                                          \
    const DecState& st = x.st;                                                                               \
    const int B = x.B, d = x.d, F = x.F, L = x.L, H = x.H, T = x.T, tmax = x.tmax;                          \
    const unsigned char* fin = x.fin;                                                                        \
    const long long cache_l = x.cache_l;                                                                     \
    const bool with_logits = x.with_logits;                                                                  \
    cudaStream_t s = x.s;                                                                                    \
    (void)F; (void)T; (void)tmax; (void)with_logits

// K5c: <= gemv_rows (64) rows, fused LayerNorm + linear kernels (gemv.cu), 8 launches per layer
static int decode_layers_fused(const StepCtx& x) {
    WSB_STEP_LOCALS;
    // <= gemv_rows (64) rows: one launch per linear layer (LayerNorm, bias, activation / residual fused), 8 per layer.
    // The residual stream's row statistics travel with it: embed -> (exact) -> qkv; out-proj -> cq;
    // cross-out -> fc1; fc2 -> next layer's qkv.
    int parts = 1;
    // folded LayerNorm (default when the checkpoint loader provided the folded tensors): the consumers read the
    // residual stream as bf16 (m->dxn, written next to the fp32 stream by its producer) and apply
    // rstd (acc - mean c1) + c2 in their epilogue; WSB_NO_FOLD=1 keeps the exact on-the-fly LayerNorm
    const bool fold = m->use_fold && m->dec[0].sqkv_wf != nullptr;
    {
        ProfScope ps(PROF_DEC_LN, 4.0 * B * d, s);
        WSB_RUN(row_stats16(m->dx, B, d, m->gv_stats, s, fold ? m->dxn : nullptr));
    }
    auto lin_ln = [&](const float* g_, const float* b_, const __nv_bfloat16* W, const float* bias, const __nv_bfloat16* Wf,
                      const float* c1, const float* c2, int N, float* out_f32, __nv_bfloat16* out_gelu) -> int {
        Gemv16Args ga;
        if (fold) {
            ga.a = m->dxn;
            ga.c1 = c1;
            ga.W = Wf;
            ga.bias = c2;
        } else {
            ga.x = m->dx;
            ga.gamma = g_;
            ga.beta = b_;
            ga.W = W;
            ga.bias = bias;
        }
        ga.stats = m->gv_stats;
        ga.stats_parts = parts;
        ga.out_f32 = out_f32;
        ga.out_bf16_gelu = out_gelu;
        ga.row_skip = fin;
        ga.fold_flag = (fold && m->fold_guard) ? m->n_active + 1 : nullptr;
        ga.M = B;
        ga.N = N;
        ga.K = d;
        ProfScope ps(PROF_DEC_GEMM, 2.0 * B * N * d, s);
        return gemv16(ga, s);
    };
    auto lin_resid = [&](const __nv_bfloat16* a_, const __nv_bfloat16* W, const float* bias, int K) -> int {
        Gemv16Args ga;
        ga.a = a_;
        ga.W = W;
        ga.bias = bias;
        ga.resid = m->dx;
        ga.xb_out = fold ? m->dxn : nullptr;
        ga.stats_out = m->gv_stats;
        ga.row_skip = fin;
        ga.M = B;
        ga.N = d;
        ga.K = K;
        parts = gemv16_parts(d, K);
        ProfScope ps(PROF_DEC_GEMM, 2.0 * B * d * K, s);
        return gemv16(ga, s);
    };
    for (int l = 0; l < L; ++l) {
        const DecLayer& e = m->dec[l];
        SplitkInput part;
        part.planes = m->dpart;
        part.splits = 1;
        part.bias = nullptr;
        WSB_RUN(lin_ln(e.ln1_g, e.ln1_b, e.sqkv_w, e.sqkv_b, e.sqkv_wf, e.sqkv_c1, e.sqkv_c2, 3 * d, m->dpart, nullptr));
        part.split_stride = static_cast<long long>(B) * 3 * d;
        WSB_RUN(decode_self_attention(nullptr, &part, d, st.k_cache + l * cache_l, st.v_cache + l * cache_l, tmax, m->step, 0,
                                      fin, m->datt, B, H, s, st.anc, st.anc_ld));
        WSB_RUN(lin_resid(m->datt, e.so_w, e.so_b, d));
        WSB_RUN(lin_ln(e.ln2_g, e.ln2_b, e.cq_w, e.cq_b, e.cq_wf, e.cq_c1, e.cq_c2, d, m->dpart, nullptr));
        part.split_stride = static_cast<long long>(B) * d;
        WSB_RUN(decode_cross_attention(nullptr, &part, d, st.cross_kv, l, L, T, fin, m->datt, B, H, s, st.kv_div));
        WSB_RUN(lin_resid(m->datt, e.co_w, e.co_b, d));
        WSB_RUN(lin_ln(e.ln3_g, e.ln3_b, e.fc1_w, e.fc1_b, e.fc1_wf, e.fc1_c1, e.fc1_c2, F, nullptr, m->dff));
        WSB_RUN(lin_resid(m->dff, e.fc2_w, e.fc2_b, F));
    }
    if (with_logits) WSB_RUN(layernorm_f32_to_bf16(m->dx, m->dec_ln_g, m->dec_ln_b, m->dxn, nullptr, B, d, s));
    return 0;
}

// K5d (opt-in): 65..256 rows, cluster split-K linear layers with folded LayerNorm (skinny.cu), 8 launches per layer
static int decode_layers_cluster(const StepCtx& x) {
    WSB_STEP_LOCALS;
    // 65..256 rows: one cluster split-K launch per linear layer (skinny.cu), LayerNorm folded into the consumers,
    // 8 launches per layer instead of 12.  The residual stream travels as fp32 (m->dx) + bf16 (m->dxn) + per-tile
    // row statistics (m->gv_stats: [parts][B][2]).
    int parts = 1;
    {
        ProfScope ps(PROF_DEC_LN, 6.0 * B * d, s);
        WSB_RUN(row_stats_any(m->dx, B, d, m->gv_stats, m->dxn, s));
    }
    auto lin_fold = [&](const __nv_bfloat16* Wf, const float* c1, const float* c2, int N, float* out_f32,
                        __nv_bfloat16* out_gelu) -> int {
        SkinnyArgs a;
        a.A = m->dxn;
        a.lda = d;
        a.W = Wf;
        a.M = B;
        a.N = N;
        a.K = d;
        a.bias = c2;
        a.c1 = c1;
        a.stats = m->gv_stats;
        a.stats_parts = parts;
        a.stats_ld = B;
        a.out_f32 = out_f32;
        a.out_bf16_gelu = out_gelu;
        a.row_skip = fin;
        ProfScope ps(PROF_DEC_GEMM, 2.0 * B * N * d, s);
        return skinny_cluster_linear(a, s);
    };
    auto lin_resid = [&](const __nv_bfloat16* a_, const __nv_bfloat16* W, const float* bias, int K) -> int {
        SkinnyArgs a;
        a.A = a_;
        a.lda = K;
        a.W = W;
        a.M = B;
        a.N = d;
        a.K = K;
        a.bias = bias;
        a.resid = m->dx;
        a.xb_out = m->dxn;
        a.stats_out = m->gv_stats;
        a.stats_out_ld = B;
        a.row_skip = fin;
        parts = d / 128;
        ProfScope ps(PROF_DEC_GEMM, 2.0 * B * d * K, s);
        return skinny_cluster_linear(a, s);
    };
    for (int l = 0; l < L; ++l) {
        const DecLayer& e = m->dec[l];
        SplitkInput part;
        part.planes = m->dpart;
        part.splits = 1;
        part.bias = nullptr;
        WSB_RUN(lin_fold(e.sqkv_wf, e.sqkv_c1, e.sqkv_c2, 3 * d, m->dpart, nullptr));
        part.split_stride = static_cast<long long>(B) * 3 * d;
        {
            ProfScope ps(PROF_DEC_SELF_ATTN, 0.0, s);
            WSB_RUN(decode_self_attention(nullptr, &part, d, st.k_cache + l * cache_l, st.v_cache + l * cache_l, tmax, m->step, 0,
                                          fin, m->datt, B, H, s, st.anc, st.anc_ld));
        }
        WSB_RUN(lin_resid(m->datt, e.so_w, e.so_b, d));
        WSB_RUN(lin_fold(e.cq_wf, e.cq_c1, e.cq_c2, d, m->dpart, nullptr));
        part.split_stride = static_cast<long long>(B) * d;
        {
            ProfScope ps(PROF_DEC_CROSS_ATTN, 4.0 * B * H * T * 64.0, s);
            WSB_RUN(decode_cross_attention(nullptr, &part, d, st.cross_kv, l, L, T, fin, m->datt, B, H, s, st.kv_div));
        }
        WSB_RUN(lin_resid(m->datt, e.co_w, e.co_b, d));
        WSB_RUN(lin_fold(e.fc1_wf, e.fc1_c1, e.fc1_c2, F, nullptr, m->dff));
        WSB_RUN(lin_resid(m->dff, e.fc2_w, e.fc2_b, F));
    }
    if (with_logits) {
        ProfScope ps(PROF_DEC_LN, 6.0 * B * d, s);
        WSB_RUN(layernorm_f32_to_bf16(m->dx, m->dec_ln_g, m->dec_ln_b, m->dxn, nullptr, B, d, s));
    }
    return 0;
}

static int decode_layers_splitk_rows(const StepCtx& x, const int B);
// K5: any batch, tcgen05 split-K GEMM + fused second phase (12 launches per layer); leaves LayerNorm(x) of the final
// decoder LayerNorm in m->dxn
static int decode_layers_splitk(const StepCtx& x) {
    // the linear layers and LayerNorms see P * B virtual rows in a prompt prefill, B rows otherwise
    return decode_layers_splitk_rows(x, x.P > 0 ? x.P * x.B : x.B);
}

static int decode_layers_splitk_rows(const StepCtx& x, const int B) {
    Model* m = x.m;
    const DecState& st = x.st;
    const int Bw = x.B, P = x.P, d = x.d, F = x.F, L = x.L, H = x.H, T = x.T, tmax = x.tmax;
    const unsigned char* fin = P > 0 ? nullptr : x.fin;
    const long long cache_l = x.cache_l;
    cudaStream_t s = x.s;
    {
        ProfScope ps(PROF_DEC_LN, 6.0 * B * d, s);
        WSB_RUN(layernorm_f32_to_bf16(m->dx, m->dec[0].ln1_g, m->dec[0].ln1_b, m->dxn, nullptr, B, d, s));
    }
    for (int l = 0; l < L; ++l) {
        const DecLayer& e = m->dec[l];
        // every LayerNorm after the first is fused into the preceding residual reduction
        const float* next_g = (l + 1 < L) ? m->dec[l + 1].ln1_g : m->dec_ln_g;
        const float* next_b = (l + 1 < L) ? m->dec[l + 1].ln1_b : m->dec_ln_b;
        SplitkInput part;
        WSB_RUN(skinny_linear(m, m->dxn, e.sqkv_w, e.sqkv_b, B, 3 * d, d, 0, nullptr, nullptr, nullptr, nullptr, nullptr, s, fin, &part,
                              m->wide_direct_bn));
        {
            ProfScope ps(PROF_DEC_SELF_ATTN, 0.0, s);
            if (P > 0)
                WSB_RUN(prefill_self_attention(&part, d, st.k_cache + l * cache_l, st.v_cache + l * cache_l, tmax, m->datt, Bw, P, H, s));
            else
                WSB_RUN(decode_self_attention(nullptr, &part, d, st.k_cache + l * cache_l, st.v_cache + l * cache_l, tmax, m->step, 0,
                                              fin, m->datt, B, H, s, st.anc, st.anc_ld));
        }
        if (m->wide_direct_resid) {
            // out-projection as one full-K launch with the in-place residual epilogue, then a plain LayerNorm
            WSB_RUN(linear(m->datt, e.so_w, e.so_b, B, d, d, GEMM_ACT_NONE, m->dx, m->dx, GEMM_OUT_F32, s, m->wide_direct_resid, PROF_DEC_GEMM));
            ProfScope ps(PROF_DEC_LN, 6.0 * B * d, s);
            WSB_RUN(layernorm_f32_to_bf16(m->dx, e.ln2_g, e.ln2_b, m->dxn, nullptr, B, d, s));
        } else {
            WSB_RUN(skinny_linear(m, m->datt, e.so_w, e.so_b, B, d, d, 0, nullptr, m->dx, e.ln2_g, e.ln2_b, m->dxn, s, fin));
        }
        WSB_RUN(skinny_linear(m, m->dxn, e.cq_w, e.cq_b, B, d, d, 0, nullptr, nullptr, nullptr, nullptr, nullptr, s, fin, &part,
                              m->wide_direct_bn ? m->wide_cq_bn : 0));
        {
            ProfScope ps(PROF_DEC_CROSS_ATTN, 4.0 * Bw * H * T * 64.0, s);   // bytes: K and V blocks, bf16
            if (P > 0)
                WSB_RUN(prefill_cross_attention(&part, d, st.cross_kv, l, L, T, m->datt, Bw, P, H, s));
            else
                WSB_RUN(decode_cross_attention(nullptr, &part, d, st.cross_kv, l, L, T, fin, m->datt, B, H, s, st.kv_div));
        }
        if (m->wide_direct_resid) {
            WSB_RUN(linear(m->datt, e.co_w, e.co_b, B, d, d, GEMM_ACT_NONE, m->dx, m->dx, GEMM_OUT_F32, s, m->wide_direct_resid, PROF_DEC_GEMM));
            ProfScope ps(PROF_DEC_LN, 6.0 * B * d, s);
            WSB_RUN(layernorm_f32_to_bf16(m->dx, e.ln3_g, e.ln3_b, m->dxn, nullptr, B, d, s));
        } else {
            WSB_RUN(skinny_linear(m, m->datt, e.co_w, e.co_b, B, d, d, 0, nullptr, m->dx, e.ln3_g, e.ln3_b, m->dxn, s, fin));
        }
        if (m->wide_direct_bn) {
            // fc1 as ONE launch: full-K tiles, bias + GELU + bf16 in the GEMM epilogue (no split-K planes, no second phase)
            ProfScope ps(PROF_DEC_GEMM, 2.0 * B * F * d, s);
            GemmArgs g;
            g.A = m->dxn;
            g.lda = d;
            g.W = e.fc1_w;
            g.M = B;
            g.N = F;
            g.K = d;
            g.bias = e.fc1_b;
            g.act = GEMM_ACT_GELU;
            g.out = m->dff;
            g.ldc = F;
            g.out_mode = GEMM_OUT_BF16;
            g.block_n = m->wide_direct_bn;
            g.row_skip = fin;
            WSB_RUN(gemm_bf16(g, s));
        } else {
            WSB_RUN(skinny_linear(m, m->dxn, e.fc1_w, e.fc1_b, B, F, d, 1, m->dff, nullptr, nullptr, nullptr, nullptr, s, fin));
        }
        WSB_RUN(skinny_linear(m, m->dff, e.fc2_w, e.fc2_b, B, d, F, 0, nullptr, m->dx, next_g, next_b, m->dxn, s, fin));
    }
    return 0;
}

static int decode_logits_tail(Model* m, const DecState& st, const __nv_bfloat16* hidden, bool first_generated, int prompt_len,
                              int max_new, const int* forced, int forced_ld, int eos_id, int pad_id, cudaStream_t s,
                              const BeamState* beam);
// one decoder position for all rows.  with_logits: project + arg-max + finalize (or, with `beam`, raw logits +
// beam bookkeeping); else prefill advance.
static int decode_step(Model* m, const DecState& st, bool with_logits, bool first_generated, int prompt_len, int max_new,
                       const int* forced, int forced_ld, int eos_id, int pad_id, cudaStream_t s,
                       const BeamState* beam = nullptr) {
    const wsb_model_config& c = m->cfg;
    const int B = st.B;
    const int d = c.d_model, F = c.ffn_dim, L = c.n_layers, H = c.n_heads, T = m->T, tmax = c.max_target_positions;
    const unsigned char* fin = forced ? nullptr : st.finished;
    struct PdlScope {
        bool prev;
        explicit PdlScope(bool on) : prev(g_use_pdl) { g_use_pdl = on; }
        ~PdlScope() { g_use_pdl = prev; }
    } pdl_scope(m->use_pdl || (m->use_gemv && B <= m->gemv_rows && c.d_model <= 1536));
    // (programmatic dependent launch: the linear-layer kernels fetch their weight tiles before the dependency wait)
    const long long cache_l = static_cast<long long>(B) * H * tmax * 64;
    const StepCtx x{m, st, B, d, F, L, H, T, tmax, fin, cache_l, with_logits, s};
    const bool folded = m->use_fold && m->dec[0].sqkv_wf != nullptr;
    if (m->use_mega && m->mega_layers != nullptr && folded && m->use_gemv && B <= m->gemv_rows && st.anc == nullptr) {
        // K5e: embedding + all decoder layers of this position in one persistent launch (mega.cu)
        MegaArgs a;
        a.layers_dev = m->mega_layers;
        a.L = L; a.d = d; a.F = F; a.H = H; a.T = T; a.tmax = tmax; a.B = B;
        a.next_token = st.next_token;
        a.step_ptr = m->step;
        a.emb = m->dec_emb;
        a.pos_emb = m->dec_pos;
        a.dx = m->dx;
        a.dxn = m->dxn;
        a.stats = m->gv_stats;
        a.proj = m->dpart;
        a.datt = m->datt;
        a.dff = m->dff;
        a.k_cache = st.k_cache;
        a.v_cache = st.v_cache;
        a.cross_kv = st.cross_kv;
        a.finished = fin;
        a.kv_div = st.kv_div;
        a.sync = m->mega_sync;
        a.fold_flag = m->fold_guard ? m->n_active + 1 : nullptr;
        a.trace = m->mega_trace;
        {
            ProfScope ps(PROF_DEC_GEMM, 0.0, s);
            WSB_RUN(decode_layers_mega(a, s));
        }
        if (with_logits) {
            ProfScope ps(PROF_DEC_LN, 6.0 * B * d, s);
            WSB_RUN(layernorm_f32_to_bf16(m->dx, m->dec_ln_g, m->dec_ln_b, m->dxn, nullptr, B, d, s));
        }
    } else {
        WSB_RUN(embed_tokens_step(st.next_token, m->step, 0, m->dec_emb, m->dec_pos, m->dx, B, d, s));
        if (m->use_gemv && B <= m->gemv_rows && d <= 1536) {
            WSB_RUN(decode_layers_fused(x));
        } else if (m->use_cluster && folded && B <= 256 && skinny_cluster_supported(B, d, d) && skinny_cluster_supported(B, F, d) &&
                   skinny_cluster_supported(B, d, F)) {
            WSB_RUN(decode_layers_cluster(x));
        } else {
            WSB_RUN(decode_layers_splitk(x));
        }
    }
    if (!with_logits) return prefill_advance(st.next_token, forced, forced_ld, m->prompt_dev, m->step, B, s);
    return decode_logits_tail(m, st, m->dxn, first_generated, prompt_len, max_new, forced, forced_ld, eos_id, pad_id, s, beam);
}

// tied output projection + logits processors + arg-max (or beam bookkeeping) on `hidden` = bf16 LayerNorm-ed rows [B][d]
static int decode_logits_tail(Model* m, const DecState& st, const __nv_bfloat16* hidden, bool first_generated, int prompt_len,
                              int max_new, const int* forced, int forced_ld, int eos_id, int pad_id, cudaStream_t s,
                              const BeamState* beam) {
    const wsb_model_config& c = m->cfg;
    const int B = st.B, d = c.d_model;
    const unsigned char* fin = forced ? nullptr : st.finished;
    if (beam) {
        // raw logits of every row (the log-softmax runs over the full vocabulary, before the suppression masks)
        GemmArgs g;
        g.A = hidden;
        g.lda = d;
        g.W = m->dec_emb;
        g.M = B;
        g.N = c.vocab_size;
        g.K = d;
        g.out = m->beam_logits;
        g.ldc = m->beam_ldv;
        g.out_mode = GEMM_OUT_F32;
        g.row_skip = fin;
        {
            ProfScope ps(PROF_DEC_LOGITS, 2.0 * B * c.vocab_size * d, s);
            WSB_RUN(gemm_bf16(g, s));
        }
        WSB_RUN(beam_step(*beam, s));
        return step_increment(m->step, s);
    }
    GemmArgs g;
    g.A = hidden;
    g.lda = d;
    g.W = m->dec_emb;
    g.M = B;
    g.N = c.vocab_size;
    g.K = d;
    g.bias = m->suppress;
    g.bias2 = first_generated ? m->begin_suppress : nullptr;
    g.out_mode = GEMM_OUT_ARGMAX;
    g.argmax_val = m->am_val;
    g.argmax_idx = m->am_idx;
    g.block_n = m->logits_bn;
    g.n_tile_list = m->logit_tiles;
    g.n_tile_count = m->am_tiles;
    {
        ProfScope ps(PROF_DEC_LOGITS, 2.0 * B * c.vocab_size * d, s);
        WSB_RUN(gemm_bf16(g, s));
    }
    return argmax_finalize(m->am_val, m->am_idx, m->am_tiles, m->tokens, max_new, prompt_len - 1, st.next_token, forced,
                           forced_ld, st.finished, m->step, m->n_active, eos_id, pad_id, B, st.row_map, s);
}

// The prompt positions of all rows in ONE pass (P * B virtual rows through the split-K path, attention kernels that serve
// the P positions of a row together), then the first generated token from the last prompt position's rows.
static int prefill_and_first_token(Model* m, const DecState& st, int prompt_len, int max_new, const int* forced, int forced_ld,
                                   int eos_id, int pad_id, cudaStream_t s) {
    const wsb_model_config& c = m->cfg;
    const int B = st.B, P = prompt_len;
    const int d = c.d_model, F = c.ffn_dim, L = c.n_layers, H = c.n_heads, T = m->T, tmax = c.max_target_positions;
    static const int kPos[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    WSB_RUN(embed_prefill(m->prompt_dev, forced, forced_ld, m->dec_emb, m->dec_pos, m->dx, B, P, d, s));
    const long long cache_l = static_cast<long long>(B) * H * tmax * 64;
    StepCtx x{m, st, B, d, F, L, H, T, tmax, nullptr, cache_l, true, s};
    x.P = P;
    WSB_RUN(decode_layers_splitk(x));
    // the decode state continues at position P - 1: its token is the last prompt token, its logits give the first new token
    WSB_CHECK_CUDA(cudaMemcpyAsync(m->step, &kPos[P - 1], sizeof(int), cudaMemcpyHostToDevice, s));
    return decode_logits_tail(m, st, m->dxn + static_cast<long long>(P - 1) * B * d, true, prompt_len, max_new, forced, forced_ld,
                              eos_id, pad_id, s, nullptr);
}

static int generate(Model* m, int B, const int* prompt, int prompt_len, int eos_id, int pad_id, int max_length,
                    const int* forced, int* tokens_out, int* n_steps, int flags, cudaStream_t s) {
    const wsb_model_config& c = m->cfg;
    WSB_REQUIRE(B >= 1 && B <= c.max_batch && B == m->last_batch, "wsb_generate must follow wsb_encode with the same batch");
    WSB_REQUIRE(prompt_len >= 1 && prompt_len <= 16, "prompt length in [1,16]");
    WSB_REQUIRE(max_length > prompt_len && max_length <= c.max_target_positions, "max_length in (prompt_len, max_target_positions]");
    const int d = c.d_model, L = c.n_layers, T = m->T, rows = B * T;
    const int max_new = max_length - prompt_len;
    m->use_pdl = (flags & 2) != 0;                      // bit1: programmatic dependent launch (GEMM weight tiles and gemv weight rows
                                                        // are fetched before the dependency wait)
    m->use_gemv = (flags & 8) == 0;                     // bit3: keep the tcgen05 split-K path for small batches too
    {
        const char* g = std::getenv("WSB_FOLD_GUARD");
        m->fold_guard = !(g && g[0] == '0');
    }
    m->use_fold = std::getenv("WSB_NO_FOLD") == nullptr && !(m->fold_guard && m->fold_disabled);
    m->use_cluster = std::getenv("WSB_CLUSTER") != nullptr;
    m->use_mega = std::getenv("WSB_MEGA") != nullptr && std::getenv("WSB_NO_MEGA") == nullptr;
    m->wide_direct_resid = 0;
    if (const char* e = std::getenv("WSB_WIDE_DIRECT_RESID")) {
        const int v = std::atoi(e);
        m->wide_direct_resid = (v == 32 || v == 64 || v == 128) ? v : 0;
    }
    m->wide_cq_bn = 32;
    if (const char* e = std::getenv("WSB_WIDE_CQ_BN")) {
        const int v = std::atoi(e);
        if (v == 32 || v == 64 || v == 128) m->wide_cq_bn = v;
    }
    m->wide_direct_bn = 128;
    if (const char* e = std::getenv("WSB_WIDE_DIRECT")) {
        const int v = std::atoi(e);
        m->wide_direct_bn = (v == 32 || v == 64 || v == 128 || v == 256) ? v : 0;
    }
    // Two persistent decode kernels on one device could starve each other (each needs every SM to make progress):
    // generate() calls that may launch them are serialised per device and drain their stream before returning.
    int cur_dev = 0;
    WSB_CHECK_CUDA(cudaGetDevice(&cur_dev));
    std::unique_lock<std::mutex> mega_lock;
    const bool mega_possible = m->use_mega && m->mega_layers != nullptr && m->use_gemv;
    if (mega_possible) mega_lock = std::unique_lock<std::mutex>(g_mega_mutex[cur_dev & 63]);
    if (const char* e = std::getenv("WSB_GEMV_ROWS")) m->gemv_rows = std::max(0, std::min(64, std::atoi(e)));
    if (const char* e = std::getenv("WSB_LADDER")) {
        std::vector<int> lv;
        for (const char* q = e; *q;) {
            const int v = std::atoi(q);
            if (v >= 1 && v <= kCompactRows && (lv.empty() || v < lv.back())) lv.push_back(v);
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (!lv.empty()) m->ladder = lv;
    }
    // cross-attention K/V of every decoder layer in one GEMM, scattered head-major:
    // cross_kv[b][layer][k|v][head][t][64]   (HF modeling_whisper.py:326-336, computed once and cached)
    {
        GemmArgs g;
        g.A = m->enc_out;
        g.lda = d;
        g.W = m->crosskv_w;
        g.M = rows;
        g.N = 2 * L * d;
        g.K = d;
        g.bias = m->crosskv_b;
        g.out = m->cross_kv;
        g.out_mode = GEMM_OUT_HEADMAJOR;
        g.rows_per_batch = T;
        ProfScope ps(PROF_CROSSKV_GEMM, 2.0 * rows * 2.0 * L * d * d, s);
        WSB_RUN(gemm_bf16(g, s));
    }
    // decode state
    std::vector<int> init_tok(B, prompt[0]);
    WSB_CHECK_CUDA(cudaMemcpyAsync(m->prompt_dev, prompt, sizeof(int) * prompt_len, cudaMemcpyHostToDevice, s));
    if (forced) {
        WSB_CHECK_CUDA(cudaMemcpy2DAsync(m->next_token, sizeof(int), forced, sizeof(int) * max_length, sizeof(int), B,
                                         cudaMemcpyDeviceToDevice, s));
    } else {
        WSB_CHECK_CUDA(cudaMemcpyAsync(m->next_token, init_tok.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
    }
    WSB_CHECK_CUDA(cudaMemsetAsync(m->step, 0, sizeof(int) * 4, s));
    WSB_CHECK_CUDA(cudaMemsetAsync(m->finished, 0, B, s));
    WSB_CHECK_CUDA(cudaMemsetAsync(m->n_active, 0, sizeof(int) * 4, s));       // [0] live rows, [1] folded-LayerNorm guard flag
    WSB_CHECK_CUDA(cudaMemcpyAsync(m->n_active, &B, sizeof(int), cudaMemcpyHostToDevice, s));
    WSB_CHECK_CUDA(cudaStreamSynchronize(s));              // init_tok / B are stack/host temporaries
    {   // pad everything: rows that never get written (early stop) must read as pad
        std::vector<int> pad(static_cast<size_t>(B) * max_new, pad_id);
        WSB_CHECK_CUDA(cudaMemcpyAsync(m->tokens, pad.data(), sizeof(int) * pad.size(), cudaMemcpyHostToDevice, s));
        WSB_CHECK_CUDA(cudaStreamSynchronize(s));
    }
    DecState st;
    st.B = B;
    st.buffer_id = 0;
    st.k_cache = m->k_cache;
    st.v_cache = m->v_cache;
    st.cross_kv = m->cross_kv;
    st.next_token = m->next_token;
    st.finished = m->finished;
    st.row_map = nullptr;
    // wide batches: the prompt positions as one pass (WSB_NO_PREFILL=1: one position at a time, as at <= 64 rows)
    const bool fused_prefill = std::getenv("WSB_NO_PREFILL") == nullptr && prompt_len >= 2 && prompt_len <= static_cast<int>(kPrefillRows) &&
                               !(m->use_gemv && B <= m->gemv_rows);
    if (fused_prefill) {
        WSB_RUN(prefill_and_first_token(m, st, prompt_len, max_new, forced, max_length, eos_id, pad_id, s));
    } else {
        for (int pos = 0; pos + 1 < prompt_len; ++pos)
            WSB_RUN(decode_step(m, st, false, false, prompt_len, max_new, forced, max_length, eos_id, pad_id, s));
    }
    if (m->use_fold && m->fold_guard && m->use_gemv && B <= m->gemv_rows && prompt_len > 1) {
        // folded-LayerNorm guard, first look: the prompt positions have run on the folded path
        WSB_CHECK_CUDA(cudaMemcpyAsync(m->pinned_active, m->n_active, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
        WSB_CHECK_CUDA(cudaStreamSynchronize(s));
        if (m->pinned_active[1] != 0) {
            m->fold_disabled = true;
            m->use_fold = false;
        }
    }
    // first generated token (begin-suppress mask active)
    if (!fused_prefill) WSB_RUN(decode_step(m, st, true, true, prompt_len, max_new, forced, max_length, eos_id, pad_id, s));
    int steps_done = 1;
    // teacher forcing bakes a caller-owned pointer into the launches: never replay those from a cached graph
    const bool use_graph = (flags & 1) != 0 && max_new > 2 && forced == nullptr;
    const bool allow_compaction = (flags & 4) == 0 && forced == nullptr;
    Model::GraphEntry* graph = nullptr;
    auto get_graph = [&](const DecState& cur, Model::GraphEntry** out) -> int {
        const auto key = std::make_tuple(cur.B, cur.buffer_id + (m->use_gemv ? 32 * m->gemv_rows + (m->use_fold ? 4096 : 0) : 16) + (m->use_cluster ? 8192 : 0) +
                                                    (m->use_pdl ? 16384 : 0) + (m->use_fold && m->fold_guard ? 32768 : 0) + (m->use_mega ? 65536 : 0) +
                                                    (std::getenv("WSB_ATTN_THREADS") ? 131072 : 0) + m->wide_direct_bn * 262144 + m->wide_direct_resid * 1000003 + m->wide_cq_bn * 7000003,
                                         cur.row_map != nullptr ? 1 : 0, max_new,
                                         prompt_len, eos_id, pad_id);
        auto it = m->graphs.find(key);
        if (it == m->graphs.end()) {
            cudaGraph_t g = nullptr;
            WSB_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            g_capturing = true;
            const long long before = g_launches.load();
            int rc = decode_step(m, cur, true, false, prompt_len, max_new, nullptr, max_length, eos_id, pad_id, s);
            const int kernels = static_cast<int>(g_launches.load() - before);
            g_launches.store(before);                   // captured, not executed
            g_capturing = false;
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc) return rc;
            WSB_CHECK_CUDA(ce);
            Model::GraphEntry entry;
            entry.kernels = kernels;
            WSB_CHECK_CUDA(cudaGraphInstantiate(&entry.exec, g, 0));
            cudaGraphDestroy(g);
            if (m->graphs.size() >= kMaxCachedGraphs) {      // folder mode with many tail batch sizes: bound the cache
                WSB_CHECK_CUDA(cudaStreamSynchronize(s));
                for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second.exec);
                m->graphs.clear();
            }
            it = m->graphs.emplace(key, entry).first;
        }
        *out = &it->second;
        return 0;
    };
    if (use_graph) WSB_RUN(get_graph(st, &graph));
    const int check_every = 4;
    while (steps_done < max_new) {
        if (use_graph) {
            ProfScope ps(PROF_DEC_GRAPH, 1.0, s);           // one replayed decoder position (all its kernels) per bracket
            WSB_CHECK_CUDA(cudaGraphLaunch(graph->exec, s));
            count_launch(graph->kernels);
        } else {
            WSB_RUN(decode_step(m, st, true, false, prompt_len, max_new, forced, max_length, eos_id, pad_id, s));
        }
        ++steps_done;
        if ((steps_done % check_every) == 0 && steps_done < max_new) {
            WSB_CHECK_CUDA(cudaMemcpyAsync(m->pinned_active, m->n_active, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
            if (mega_possible)
                WSB_CHECK_CUDA(cudaMemcpyAsync(m->pinned_active + 2, m->mega_sync + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
            WSB_CHECK_CUDA(cudaStreamSynchronize(s));
            if (mega_possible && m->pinned_active[2] != 0) {
                cudaMemsetAsync(m->mega_sync, 0, sizeof(unsigned int) * 64, s);
                set_last_error("persistent decode kernel: grid barrier watchdog fired (another kernel is holding SMs of this device?)");
                return 6;
            }
            if (m->pinned_active[1] != 0 && m->use_fold && m->fold_guard) {
                // folded-LayerNorm guard: some live row's mean dominates its spread -> exact LayerNorm from here on
                m->fold_disabled = true;
                m->use_fold = false;
                if (use_graph) WSB_RUN(get_graph(st, &graph));
            }
            if (forced) continue;
            const int active = m->pinned_active[0];
            if (active <= 0) break;
            // batch compaction ladder: every move goes to the other buffer set (main <-> alternate); the target is
            // the smallest level that still holds the active rows
            int target = 0;
            if (allow_compaction && max_new - steps_done >= 16) {
                for (int lv : m->ladder)
                    if (lv < st.B && active <= lv) target = lv;
            }
            if (target > 0) {
                DecState nx;
                nx.B = target;
                nx.buffer_id = 1 - st.buffer_id;
                const bool to_alt = nx.buffer_id == 1;
                nx.k_cache = to_alt ? m->k_cache_alt : m->k_cache;
                nx.v_cache = to_alt ? m->v_cache_alt : m->v_cache;
                nx.cross_kv = to_alt ? m->cross_kv_alt : m->cross_kv;
                nx.next_token = to_alt ? m->next_token_alt : m->next_token;
                nx.finished = to_alt ? m->finished_alt : m->finished;
                int* map_dst = to_alt ? m->row_map_alt : m->row_map_main;
                nx.row_map = map_dst;
                CompactArgs ca;
                ca.fin_src = st.finished;
                ca.b_src = st.B;
                ca.map_src = st.row_map;
                ca.tok_src = st.next_token;
                ca.b_dst = target;
                ca.active_idx = m->active_idx;
                ca.map_dst = map_dst;
                ca.tok_dst = nx.next_token;
                ca.fin_dst = nx.finished;
                ca.k_src = st.k_cache;
                ca.v_src = st.v_cache;
                ca.cross_src = st.cross_kv;
                ca.k_dst = nx.k_cache;
                ca.v_dst = nx.v_cache;
                ca.cross_dst = nx.cross_kv;
                ca.step_ptr = m->step;
                ca.n_heads = c.n_heads;
                ca.n_layers = L;
                ca.t_max = c.max_target_positions;
                ca.cross_row_elems = static_cast<long long>(L) * 2 * c.n_heads * T * 64;
                {
                    ProfScope ps(PROF_DEC_COMPACT, 1.0, s);
                    WSB_RUN(compact_decode_state(ca, s));
                }
                st = nx;
                if (use_graph) WSB_RUN(get_graph(st, &graph));
            }
        }
    }
    WSB_CHECK_CUDA(cudaMemcpyAsync(tokens_out, m->tokens, sizeof(int) * static_cast<size_t>(B) * max_new,
                                   cudaMemcpyDeviceToDevice, s));
    if (mega_possible) {
        WSB_CHECK_CUDA(cudaMemcpyAsync(m->pinned_active + 2, m->mega_sync + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
        WSB_CHECK_CUDA(cudaStreamSynchronize(s));          // nothing of this call is still running when the device lock drops
        if (m->pinned_active[2] != 0) {
            cudaMemsetAsync(m->mega_sync, 0, sizeof(unsigned int) * 64, s);
            set_last_error("persistent decode kernel: grid barrier watchdog fired (another kernel is holding SMs of this device?)");
            return 6;
        }
    }
    if (n_steps) *n_steps = steps_done;
    return 0;
}

// Beam search over `B` windows x `nb` beams (rows = B*nb <= max_batch): the reference's default decode mode
// (HF generate(num_beams=4, length_penalty), reference model.py:409, 614, 662).  See beam.cu.
static int generate_beam(Model* m, int B, int nb, const int* prompt, int prompt_len, int eos_id, int pad_id, int max_length,
                         float length_penalty, int* tokens_out, float* scores_out, int* n_steps, int flags, cudaStream_t s) {
    const wsb_model_config& c = m->cfg;
    WSB_REQUIRE(nb >= 1 && nb <= 4, "num_beams in [1,4]");
    const int R = B * nb;
    WSB_REQUIRE(B >= 1 && R <= c.max_batch && B == m->last_batch, "wsb_generate_beam must follow wsb_encode with batch*num_beams <= max_batch");
    WSB_REQUIRE(prompt_len >= 1 && prompt_len <= 16, "prompt length in [1,16]");
    WSB_REQUIRE(max_length > prompt_len && max_length <= c.max_target_positions, "max_length in (prompt_len, max_target_positions]");
    const int d = c.d_model, L = c.n_layers, T = m->T, rows = B * T;
    const int max_new = max_length - prompt_len;
    m->use_pdl = false;
    {
        const char* g = std::getenv("WSB_FOLD_GUARD");
        m->fold_guard = !(g && g[0] == '0');
    }
    m->use_fold = std::getenv("WSB_NO_FOLD") == nullptr && !(m->fold_guard && m->fold_disabled);
    if (!m->beam_ws) {
        m->beam_ldv = (c.vocab_size + 7) & ~7;
        const size_t logits_bytes = (static_cast<size_t>(c.max_batch) * m->beam_ldv * sizeof(float) + 255) & ~size_t(255);
        WSB_CHECK_CUDA(cudaMalloc(&m->beam_ws, logits_bytes + beam_state_bytes(c.max_batch, c.max_target_positions)));
        m->beam_logits = reinterpret_cast<float*>(m->beam_ws);
        beam_state_carve(&m->beam, m->beam_ws + logits_bytes, c.max_batch, c.max_target_positions);
    }
    {   // cross-attention K/V once per window (shared by its beams)
        GemmArgs g;
        g.A = m->enc_out;
        g.lda = d;
        g.W = m->crosskv_w;
        g.M = rows;
        g.N = 2 * L * d;
        g.K = d;
        g.bias = m->crosskv_b;
        g.out = m->cross_kv;
        g.out_mode = GEMM_OUT_HEADMAJOR;
        g.rows_per_batch = T;
        ProfScope ps(PROF_CROSSKV_GEMM, 2.0 * rows * 2.0 * L * d * d, s);
        WSB_RUN(gemm_bf16(g, s));
    }
    BeamState& bs = m->beam;
    bs.B = B;
    bs.nb = nb;
    bs.K = 2 * nb;
    bs.V = c.vocab_size;
    bs.ldv = m->beam_ldv;
    bs.max_length = max_length;
    bs.prompt_len = prompt_len;
    bs.eos_id = eos_id;
    bs.pad_id = pad_id;
    bs.logits = m->beam_logits;
    bs.suppress = m->suppress;
    bs.begin_suppress = m->begin_suppress;
    bs.next_token = m->next_token;
    bs.finished = m->finished;
    bs.step_ptr = m->step;
    bs.n_active = m->n_active;
    WSB_CHECK_CUDA(cudaMemcpyAsync(m->prompt_dev, prompt, sizeof(int) * prompt_len, cudaMemcpyHostToDevice, s));
    WSB_CHECK_CUDA(cudaMemsetAsync(m->step, 0, sizeof(int) * 4, s));
    WSB_CHECK_CUDA(cudaMemsetAsync(m->n_active, 0, sizeof(int) * 4, s));
    WSB_CHECK_CUDA(cudaMemcpyAsync(m->n_active, &B, sizeof(int), cudaMemcpyHostToDevice, s));
    WSB_RUN(beam_set_length_penalty(bs, length_penalty, max_length, s));      // synchronises: prompt / B are temporaries
    WSB_RUN(beam_init(bs, m->prompt_dev, s));
    DecState st;
    st.B = R;
    st.buffer_id = 0;
    st.k_cache = m->k_cache;
    st.v_cache = m->v_cache;
    st.cross_kv = m->cross_kv;
    st.next_token = m->next_token;
    st.finished = m->finished;
    st.row_map = nullptr;
    st.kv_div = nb;
    st.anc = bs.anc;
    st.anc_ld = bs.seq_ld;
    for (int pos = 0; pos + 1 < prompt_len; ++pos)
        WSB_RUN(decode_step(m, st, false, false, prompt_len, max_new, nullptr, max_length, eos_id, pad_id, s));
    WSB_RUN(decode_step(m, st, true, true, prompt_len, max_new, nullptr, max_length, eos_id, pad_id, s, &bs));
    int steps_done = 1;
    const bool use_graph = (flags & 1) != 0 && max_new > 2;
    Model::GraphEntry* graph = nullptr;
    // (the length penalty lives in a device table, not in the launches)
    auto get_graph = [&](Model::GraphEntry** out) -> int {
        const auto key = std::make_tuple(B, nb, max_length, prompt_len, eos_id, pad_id,
                                         (m->use_fold ? 1 : 0) | (m->use_gemv ? 2 : 0) | (m->fold_guard ? 4 : 0) | (m->gemv_rows << 3));
        auto it = m->beam_graphs.find(key);
        if (it == m->beam_graphs.end()) {
            cudaGraph_t g = nullptr;
            WSB_CHECK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            g_capturing = true;
            const long long before = g_launches.load();
            int rc = decode_step(m, st, true, false, prompt_len, max_new, nullptr, max_length, eos_id, pad_id, s, &bs);
            const int kernels = static_cast<int>(g_launches.load() - before);
            g_launches.store(before);
            g_capturing = false;
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc) return rc;
            WSB_CHECK_CUDA(ce);
            Model::GraphEntry entry;
            entry.kernels = kernels;
            WSB_CHECK_CUDA(cudaGraphInstantiate(&entry.exec, g, 0));
            cudaGraphDestroy(g);
            it = m->beam_graphs.emplace(key, entry).first;
        }
        *out = &it->second;
        return 0;
    };
    if (use_graph) WSB_RUN(get_graph(&graph));
    const int check_every = 4;
    while (steps_done < max_new) {
        if (use_graph) {
            ProfScope ps(PROF_DEC_GRAPH, 1.0, s);
            WSB_CHECK_CUDA(cudaGraphLaunch(graph->exec, s));
            count_launch(graph->kernels);
        } else {
            WSB_RUN(decode_step(m, st, true, false, prompt_len, max_new, nullptr, max_length, eos_id, pad_id, s, &bs));
        }
        ++steps_done;
        if ((steps_done % check_every) == 0 && steps_done < max_new) {
            WSB_CHECK_CUDA(cudaMemcpyAsync(m->pinned_active, m->n_active, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
            WSB_CHECK_CUDA(cudaStreamSynchronize(s));
            if (m->pinned_active[0] <= 0) break;
            if (m->pinned_active[1] != 0 && m->use_fold && m->fold_guard) {   // folded-LayerNorm guard (see generate)
                m->fold_disabled = true;
                m->use_fold = false;
                if (use_graph) WSB_RUN(get_graph(&graph));
            }
        }
    }
    WSB_RUN(beam_output(bs, tokens_out, scores_out, max_new, s));
    if (n_steps) *n_steps = steps_done;
    return 0;
}

}  // namespace wsb

// =============================================================================================== C ABI
using namespace wsb;

struct wsb_logmel_plan {
    LogmelPlan* impl;
};
struct wsb_model {
    Model* impl;
};

extern "C" {

int wsb_abi_version(void) { return WSB_ABI_VERSION; }
const char* wsb_last_error(void) { return get_last_error(); }
long long wsb_launch_count(int reset) {
    long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}
int wsb_set_sm_reserve(int n_sms) {
    const int prev = g_sm_reserve;
    g_sm_reserve = n_sms < 0 ? 0 : n_sms;
    return prev;
}

int wsb_logmel_plan_create(int n_fft, int hop, int clip_len, int n_cols, const float* mel_filters, int n_freq,
                           wsb_logmel_plan** plan) {
    WSB_REQUIRE(plan != nullptr && mel_filters != nullptr, "null argument");
    LogmelPlan* impl = nullptr;
    int rc = logmel_plan_create(n_fft, hop, clip_len, n_cols, mel_filters, n_freq, &impl);
    if (rc) return rc;
    *plan = new wsb_logmel_plan{impl};
    return 0;
}
void wsb_logmel_plan_destroy(wsb_logmel_plan* plan) {
    if (!plan) return;
    logmel_plan_destroy(plan->impl);
    delete plan;
}
int wsb_logmel_run(const wsb_logmel_plan* plan, const float* audio_dev, const int64_t* windows_dev, int n_windows,
                   float* features_dev, void* stream) {
    WSB_REQUIRE(plan != nullptr, "null plan");
    return logmel_run(plan->impl, audio_dev, reinterpret_cast<const long long*>(windows_dev), n_windows, features_dev,
                      static_cast<cudaStream_t>(stream));
}

int wsb_model_create(const wsb_model_config* cfg, const char* const* names, const void* const* tensors_dev,
                     int n_tensors, wsb_model** model) {
    WSB_REQUIRE(cfg && names && tensors_dev && model, "null argument");
    Model* impl = nullptr;
    int rc = model_create(cfg, names, tensors_dev, n_tensors, &impl);
    if (rc) return rc;
    *model = new wsb_model{impl};
    return 0;
}
void wsb_model_destroy(wsb_model* model) {
    if (!model) return;
    model_destroy(model->impl);
    delete model;
}
size_t wsb_workspace_bytes_for(const wsb_model_config* cfg) {
    if (!cfg || cfg->max_batch < 1) return 0;
    Model m;
    m.cfg = *cfg;
    m.T = cfg->n_cols / 2;
    m.am_tiles = gemm_n_tiles(cfg->vocab_size, 32);        // upper bound: every vocabulary tile of the narrowest block
    model_layout(&m, false);
    return m.ws_bytes;
}
int wsb_mega_trace(wsb_model* model, unsigned long long* out_host, int n) {
    WSB_REQUIRE(model != nullptr && model->impl != nullptr, "null model");
    Model* m = model->impl;
    const int cap = 2 * (2 + 10 * m->cfg.n_layers) + 2 + 64;
    if (out_host == nullptr) {                              // arm (n != 0) or disarm (n == 0) the trace buffer
        if (n != 0 && m->mega_trace == nullptr) {
            WSB_CHECK_CUDA(cudaMalloc(&m->mega_trace, sizeof(unsigned long long) * cap));
            WSB_CHECK_CUDA(cudaMemset(m->mega_trace, 0, sizeof(unsigned long long) * cap));
            for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second.exec);     // the pointer is baked into captured launches
            m->graphs.clear();
        }
        return cap;
    }
    WSB_REQUIRE(m->mega_trace != nullptr, "trace not armed");
    WSB_CHECK_CUDA(cudaDeviceSynchronize());
    WSB_CHECK_CUDA(cudaMemcpy(out_host, m->mega_trace, sizeof(unsigned long long) * std::min(n, cap), cudaMemcpyDeviceToHost));
    WSB_CHECK_CUDA(cudaMemset(m->mega_trace + (cap - 64), 0, sizeof(unsigned long long) * 64));      // stage accumulators restart
    return std::min(n, cap);
}
int wsb_model_fold_fallback(const wsb_model* model) { return (model && model->impl && model->impl->fold_disabled) ? 1 : 0; }
size_t wsb_model_workspace_bytes(const wsb_model* model) { return model ? model->impl->ws_bytes : 0; }

int wsb_encode(wsb_model* model, const float* features_dev, int batch, float* hidden_f32_dev, void* stream) {
    WSB_REQUIRE(model != nullptr, "null model");
    return encode(model->impl, features_dev, batch, hidden_f32_dev, static_cast<cudaStream_t>(stream));
}
int wsb_generate(wsb_model* model, int batch, const int32_t* prompt, int prompt_len, int eos_id, int pad_id,
                 int max_length, const int32_t* forced_dev, int32_t* tokens_dev, int* n_steps, int flags, void* stream) {
    WSB_REQUIRE(model != nullptr && prompt != nullptr && tokens_dev != nullptr, "null argument");
    return generate(model->impl, batch, prompt, prompt_len, eos_id, pad_id, max_length, forced_dev, tokens_dev, n_steps,
                    flags, static_cast<cudaStream_t>(stream));
}

int wsb_generate_beam(wsb_model* model, int batch, int num_beams, const int32_t* prompt, int prompt_len, int eos_id,
                      int pad_id, int max_length, float length_penalty, int32_t* tokens_dev, float* scores_dev, int* n_steps,
                      int flags, void* stream) {
    WSB_REQUIRE(model != nullptr && prompt != nullptr && tokens_dev != nullptr, "null argument");
    return generate_beam(model->impl, batch, num_beams, prompt, prompt_len, eos_id, pad_id, max_length, length_penalty,
                         tokens_dev, scores_dev, n_steps, flags, static_cast<cudaStream_t>(stream));
}

int wsb_beam_selftest(int batch, int num_beams, int vocab, int n_steps, const float* logits_dev, const float* suppress_dev,
                      const int32_t* prompt, int prompt_len, int eos_id, int pad_id, int max_length, float length_penalty,
                      int32_t* tokens_dev, float* scores_dev, int32_t* parents_dev, int32_t* next_tokens_dev, void* stream) {
    WSB_REQUIRE(logits_dev && suppress_dev && prompt && tokens_dev, "null argument");
    WSB_REQUIRE(num_beams >= 1 && num_beams <= 4 && batch >= 1, "num_beams in [1,4]");
    WSB_REQUIRE(prompt_len >= 1 && prompt_len <= 16 && max_length > prompt_len && max_length <= 512, "bad lengths");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int R = batch * num_beams, ld = max_length;
    char* ws = nullptr;
    const size_t extra = 4096 + static_cast<size_t>(R) * 8;
    WSB_CHECK_CUDA(cudaMalloc(&ws, beam_state_bytes(R, ld) + extra));
    struct Free {
        char* p;
        ~Free() { cudaFree(p); }
    } guard{ws};
    BeamState bs;
    int* step = reinterpret_cast<int*>(ws);
    int* n_active = step + 4;
    int* prompt_dev = step + 8;
    bs.next_token = reinterpret_cast<int*>(ws + 1024);
    bs.finished = reinterpret_cast<unsigned char*>(ws + 1024 + static_cast<size_t>(R) * 4);
    beam_state_carve(&bs, ws + ((extra + 255) & ~size_t(255)), R, ld);
    bs.B = batch;
    bs.nb = num_beams;
    bs.K = 2 * num_beams;
    bs.V = vocab;
    bs.ldv = vocab;
    bs.max_length = max_length;
    bs.prompt_len = prompt_len;
    bs.eos_id = eos_id;
    bs.pad_id = pad_id;
    bs.suppress = suppress_dev;
    bs.begin_suppress = nullptr;
    bs.step_ptr = step;
    bs.n_active = n_active;
    const int start = prompt_len - 1;
    WSB_CHECK_CUDA(cudaMemcpyAsync(prompt_dev, prompt, sizeof(int) * prompt_len, cudaMemcpyHostToDevice, s));
    WSB_CHECK_CUDA(cudaMemcpyAsync(step, &start, sizeof(int), cudaMemcpyHostToDevice, s));
    WSB_CHECK_CUDA(cudaMemcpyAsync(n_active, &batch, sizeof(int), cudaMemcpyHostToDevice, s));
    WSB_RUN(beam_set_length_penalty(bs, length_penalty, max_length, s));
    WSB_RUN(beam_init(bs, prompt_dev, s));
    for (int t = 0; t < n_steps && prompt_len + t < max_length; ++t) {
        bs.logits = logits_dev + static_cast<size_t>(t) * R * vocab;
        WSB_RUN(beam_step(bs, s));
        const int pos = start + t, nbuf = (pos & 1) ^ 1;
        if (parents_dev)
            WSB_CHECK_CUDA(cudaMemcpy2DAsync(parents_dev + static_cast<size_t>(t) * R, sizeof(int),
                                             bs.anc + static_cast<size_t>(nbuf) * R * ld + pos, sizeof(int) * ld, sizeof(int), R,
                                             cudaMemcpyDeviceToDevice, s));
        if (next_tokens_dev)
            WSB_CHECK_CUDA(cudaMemcpyAsync(next_tokens_dev + static_cast<size_t>(t) * R, bs.next_token, sizeof(int) * R,
                                           cudaMemcpyDeviceToDevice, s));
        WSB_RUN(step_increment(step, s));
    }
    WSB_RUN(beam_output(bs, tokens_dev, scores_dev, max_length - prompt_len, s));
    WSB_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int wsb_profile_enable(int enable) {
    g_prof.enabled = enable != 0;
    if (enable) {
        g_prof.collect();
        for (int i = 0; i < PROF_NCAT; ++i) {
            g_prof.ms[i] = 0;
            g_prof.work[i] = 0;
            g_prof.count[i] = 0;
        }
    }
    return 0;
}
int wsb_profile_read(int category, double* ms, long long* launches, double* work) {
    WSB_REQUIRE(category >= 0 && category < PROF_NCAT, "bad profile category");
    g_prof.collect();                                   // caller has synchronised the stream
    if (ms) *ms = g_prof.ms[category];
    if (launches) *launches = g_prof.count[category];
    if (work) *work = g_prof.work[category];
    return 0;
}

int wsb_gemm_bf16(const void* a_dev, const void* w_dev, int M, int N, int K, const float* bias_dev, int gelu,
                  const float* resid_dev, void* c_dev, int out_f32, int block_n, void* stream) {
    return linear(static_cast<const __nv_bfloat16*>(a_dev), static_cast<const __nv_bfloat16*>(w_dev), bias_dev, M, N, K,
                  gelu ? GEMM_ACT_GELU : GEMM_ACT_NONE, resid_dev, c_dev, out_f32 ? GEMM_OUT_F32 : GEMM_OUT_BF16,
                  static_cast<cudaStream_t>(stream), block_n);
}
int wsb_layernorm(const float* x_dev, const float* gamma_dev, const float* beta_dev, void* out_bf16_dev,
                  float* out_f32_dev, int rows, int d, void* stream) {
    return layernorm_f32_to_bf16(x_dev, gamma_dev, beta_dev, static_cast<__nv_bfloat16*>(out_bf16_dev), out_f32_dev, rows,
                                 d, static_cast<cudaStream_t>(stream));
}
int wsb_gemv16(const float* x_f32_dev, const float* gamma_dev, const float* beta_dev, const void* a_bf16_dev, const void* w_dev,
               const float* bias_dev, int M, int N, int K, int out_mode, void* out_dev, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Gemv16Args g;
    float* stats = nullptr;
    struct Free {
        float*& p;
        ~Free() { cudaFree(p); }
    } guard{stats};
    WSB_CHECK_CUDA(cudaMalloc(&stats, sizeof(float) * 640 * 128 * 2));
    __nv_bfloat16* xb = nullptr;
    struct FreeB {
        __nv_bfloat16*& p;
        ~FreeB() { cudaFree(p); }
    } guard_b{xb};
    g.a = static_cast<const __nv_bfloat16*>(a_bf16_dev);
    if (x_f32_dev && gamma_dev && !beta_dev) {            // folded LayerNorm: gamma_dev = c1, w_dev = bf16(W o gamma), bias_dev = c2
        WSB_REQUIRE(M >= 1 && M <= gemv16_max_rows() && K >= 32, "gemv16 shape");
        WSB_CHECK_CUDA(cudaMalloc(&xb, sizeof(__nv_bfloat16) * static_cast<size_t>(M) * K));
        WSB_RUN(row_stats16(x_f32_dev, M, K, stats, s, xb));
        g.a = xb;
        g.c1 = gamma_dev;
        g.stats = stats;
        g.stats_parts = 1;
    } else if (x_f32_dev) {                               // stand-alone use: exact row statistics, one part
        WSB_RUN(row_stats16(x_f32_dev, M, K, stats, s));
        g.x = x_f32_dev;
        g.stats = stats;
        g.stats_parts = 1;
        g.gamma = gamma_dev;
        g.beta = beta_dev;
    }
    g.W = static_cast<const __nv_bfloat16*>(w_dev);
    g.bias = bias_dev;
    if (out_mode == 0) g.out_f32 = static_cast<float*>(out_dev);
    else if (out_mode == 1) g.out_bf16_gelu = static_cast<__nv_bfloat16*>(out_dev);
    else {
        g.resid = static_cast<float*>(out_dev);
        g.stats_out = stats + 640 * 128;
    }
    g.M = M;
    g.N = N;
    g.K = K;
    WSB_RUN(gemv16(g, s));
    WSB_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}
int wsb_skinny_linear(const float* x_f32_dev, const float* c1_dev, const void* a_bf16_dev, const void* w_dev, const float* bias_dev,
                      int M, int N, int K, int out_mode, int splits, void* out_dev, void* xb_out_dev, float* stats_out_dev,
                      void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SkinnyArgs g;
    float* stats = nullptr;
    __nv_bfloat16* xb = nullptr;
    struct Free {
        float*& a;
        __nv_bfloat16*& b;
        ~Free() {
            cudaFree(a);
            cudaFree(b);
        }
    } guard{stats, xb};
    if (x_f32_dev) {                                      // folded LayerNorm: exact row statistics (one part) + bf16 rows
        WSB_REQUIRE(M >= 1 && K >= 1 && c1_dev && out_mode <= 1, "wsb_skinny_linear: folded form needs c1 and out_mode 0 / 1");
        WSB_CHECK_CUDA(cudaMalloc(&stats, sizeof(float) * 2 * M));
        WSB_CHECK_CUDA(cudaMalloc(&xb, sizeof(__nv_bfloat16) * static_cast<size_t>(M) * K));
        WSB_RUN(row_stats_any(x_f32_dev, M, K, stats, xb, s));
        g.A = xb;
        g.c1 = c1_dev;
        g.stats = stats;
        g.stats_parts = 1;
        g.stats_ld = M;
    } else {
        WSB_REQUIRE(a_bf16_dev && out_mode == 2, "wsb_skinny_linear: bf16 activations feed the residual form (out_mode 2)");
        g.A = static_cast<const __nv_bfloat16*>(a_bf16_dev);
    }
    g.lda = K;
    g.W = static_cast<const __nv_bfloat16*>(w_dev);
    g.bias = bias_dev;
    g.M = M;
    g.N = N;
    g.K = K;
    g.splits = splits;
    if (out_mode == 0) g.out_f32 = static_cast<float*>(out_dev);
    else if (out_mode == 1) g.out_bf16_gelu = static_cast<__nv_bfloat16*>(out_dev);
    else {
        g.resid = static_cast<float*>(out_dev);
        g.xb_out = static_cast<__nv_bfloat16*>(xb_out_dev);
        g.stats_out = stats_out_dev;
        g.stats_out_ld = M;
    }
    WSB_RUN(skinny_cluster_linear(g, s));
    WSB_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}
int wsb_gemv16_bench(int M, int N, int K, int mode, int iters, int weight_copies, float* us_per_launch) {
    // diagnostics: steady-state time of one skinny-linear launch; the weights rotate through `weight_copies`
    // buffers so that they come from HBM, not L2.  mode 0: LN -> fp32, 1: LN -> GELU bf16, 2: bf16 -> residual, 3 / 4: folded LN -> fp32 / GELU bf16
    WSB_REQUIRE(M >= 1 && M <= gemv16_max_rows() && iters >= 1 && weight_copies >= 1 && us_per_launch, "bad arguments");
    const size_t wn = static_cast<size_t>(N) * K;
    __nv_bfloat16 *w = nullptr, *a = nullptr;
    float *x = nullptr, *gb = nullptr, *stats = nullptr, *out = nullptr;
    struct Free {
        void** p[6];
        ~Free() { for (auto q : p) cudaFree(*q); }
    } guard{{reinterpret_cast<void**>(&w), reinterpret_cast<void**>(&a), reinterpret_cast<void**>(&x), reinterpret_cast<void**>(&gb),
             reinterpret_cast<void**>(&stats), reinterpret_cast<void**>(&out)}};
    WSB_CHECK_CUDA(cudaMalloc(&w, wn * 2 * weight_copies));
    WSB_CHECK_CUDA(cudaMalloc(&a, static_cast<size_t>(64) * K * 2));
    WSB_CHECK_CUDA(cudaMalloc(&x, static_cast<size_t>(64) * K * 4));
    WSB_CHECK_CUDA(cudaMalloc(&gb, static_cast<size_t>(K) * 8 + static_cast<size_t>(N) * 4));
    WSB_CHECK_CUDA(cudaMalloc(&stats, sizeof(float) * 640 * 128 * 2));
    WSB_CHECK_CUDA(cudaMalloc(&out, static_cast<size_t>(64) * N * 4));
    WSB_CHECK_CUDA(cudaMemset(w, 0, wn * 2 * weight_copies));
    WSB_CHECK_CUDA(cudaMemset(a, 0, static_cast<size_t>(64) * K * 2));
    WSB_CHECK_CUDA(cudaMemset(x, 0, static_cast<size_t>(64) * K * 4));
    WSB_CHECK_CUDA(cudaMemset(gb, 0, static_cast<size_t>(K) * 8 + static_cast<size_t>(N) * 4));
    WSB_CHECK_CUDA(cudaMemset(stats, 0, sizeof(float) * 640 * 128 * 2));
    WSB_CHECK_CUDA(cudaMemset(out, 0, static_cast<size_t>(64) * N * 4));
    cudaStream_t s = nullptr;
    cudaEvent_t e0, e1;
    WSB_CHECK_CUDA(cudaEventCreate(&e0));
    WSB_CHECK_CUDA(cudaEventCreate(&e1));
    Gemv16Args g;
    if (mode <= 1) {
        g.x = x;
        g.stats = stats;
        g.stats_parts = gemv16_parts(K, K);
        g.gamma = gb;
        g.beta = gb + K;
    } else if (mode >= 3) {                             // folded LayerNorm: 3 -> fp32, 4 -> GELU bf16
        g.a = a;
        g.c1 = gb + 2 * K;
        g.stats = stats;
        g.stats_parts = gemv16_parts(K, K);
    } else {
        g.a = a;
    }
    g.bias = gb + 2 * K;
    if (mode == 0 || mode == 3) g.out_f32 = out;
    else if (mode == 1 || mode == 4) g.out_bf16_gelu = reinterpret_cast<__nv_bfloat16*>(out);
    else {
        g.resid = out;
        g.stats_out = stats + 640 * 128;
    }
    g.M = M;
    g.N = N;
    g.K = K;
    for (int it = -3; it < iters; ++it) {
        if (it == 0) WSB_CHECK_CUDA(cudaEventRecord(e0, s));
        g.W = w + wn * ((it + 3) % weight_copies);
        WSB_RUN(gemv16(g, s));
    }
    WSB_CHECK_CUDA(cudaEventRecord(e1, s));
    WSB_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.0f;
    WSB_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *us_per_launch = ms * 1000.0f / iters;
    return 0;
}
int wsb_encoder_attention(const void* qkv_dev, void* out_dev, int batch, int T, int n_heads, void* stream) {
    return encoder_attention(static_cast<const __nv_bfloat16*>(qkv_dev), static_cast<__nv_bfloat16*>(out_dev), batch, T,
                             n_heads, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
