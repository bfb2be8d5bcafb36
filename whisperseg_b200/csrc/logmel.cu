// K1 -- fused log-mel front-end for a batch of sliding windows.
//
// Replaces, per window, the chain the reference runs on the CPU (model.py:146-165 calling HF
// WhisperFeatureExtractor._torch_extract_fbank_features, feature_extraction_whisper.py:135-164):
//   zero-padded clip -> reflect-centred framing -> periodic Hann -> rFFT -> |.|^2 -> slaney mel
//   -> log10(clamp 1e-10) -> max(., window_max - 8) -> (x+4)/4 -> first 1000 columns
//   (+ right-padding with the window minimum when the clip has fewer than 1000 frames).
//
// One thread-block CLUSTER per window.  Each CTA of the cluster owns a contiguous slice of
// frames: it stages the slice's samples in shared memory once (float4 global loads, every
// sample is read from HBM exactly once per CTA that needs it), runs a shared-memory Stockham
// real-FFT per frame, applies the sparse (CSR) mel filterbank and log10, and keeps its
// [80 x frames] tile on chip.  The window-global max (and min) is reduced across the cluster
// through distributed shared memory, then the clamp + affine is applied while the tile is
// written out -- so HBM traffic is exactly samples-in + features-out.
#include "common.cuh"
#include "wsb_internal.h"

#include <cooperative_groups.h>
#include <math.h>
#include <vector>

namespace cg = cooperative_groups;

namespace wsb {

constexpr int kLogmelThreads = 256;
constexpr int kMels = 80;
constexpr int kMaxCluster = 8;

struct LogmelPlan {
    int n_fft, hop, clip_len, n_frames, n_cols, log2m;
    int cluster, frames_per_cta, group_threads, n_groups;
    int span_floats, tile_stride;
    size_t smem_bytes;
    float* hann = nullptr;          // [n_fft]
    float2* tw = nullptr;           // [n_fft/2]  exp(-2 pi i j / n_fft)
    int* mel_start = nullptr;       // [80]
    int* mel_cnt = nullptr;         // [80]
    int* mel_off = nullptr;         // [80]
    float* mel_w = nullptr;         // [nnz]
    int nnz = 0;
};

struct LogmelParams {
    const float* audio;
    const long long* win;           // [n_win][3] = start, lo, hi  (audio sample indices; valid iff lo <= a < hi)
    float* out;                     // [n_win][80][n_cols]
    const float* hann;
    const float2* tw;
    const int* mel_start;
    const int* mel_cnt;
    const int* mel_off;
    const float* mel_w;
    int n_fft, hop, clip_len, n_frames, n_cols, log2m;
    int frames_per_cta, group_threads, n_groups, span_floats, tile_stride;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// exp(-2 pi i idx / n_fft) for idx in [0, n_fft) from the half table
__device__ __forceinline__ float2 twiddle(const float2* __restrict__ tw, int idx, int m) {
    float2 t = tw[idx >= m ? idx - m : idx];
    return idx >= m ? make_float2(-t.x, -t.y) : t;
}

// barrier among the G threads that cooperate on one frame (a warp, or a named barrier per group), so
// the frame groups of a CTA run their FFT passes independently of each other
__device__ __forceinline__ void group_sync(int g, int G) {
    if (G == 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(G) : "memory");
    }
}

__global__ void __launch_bounds__(kLogmelThreads) logmel_kernel(const LogmelParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int csize = static_cast<int>(cluster.num_blocks());
    const int w = blockIdx.x / csize;
    const int tid = threadIdx.x;
    const int M = p.n_fft >> 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_samples = reinterpret_cast<float*>(smem_raw);                       // span_floats
    float* s_tile = s_samples + p.span_floats;                                   // 80 * tile_stride
    float2* s_tw = reinterpret_cast<float2*>(s_tile + kMels * p.tile_stride);    // M
    float2* s_fft = s_tw + M;                                                    // n_groups * 2 * M
    __shared__ float s_red[2][kLogmelThreads / 32];
    __shared__ float s_cluster_red[2];                                           // this CTA's {max, min}, read by peers

    const int f0 = rank * p.frames_per_cta;
    const int f1 = min(p.n_frames, f0 + p.frames_per_cta);
    const int nf = max(0, f1 - f0);

    const long long start = p.win[3 * w + 0];
    const long long lo = p.win[3 * w + 1];
    const long long hi = p.win[3 * w + 2];

    // ---- stage twiddles and this CTA's sample span -------------------------------------------
    for (int i = tid; i < M; i += kLogmelThreads) s_tw[i] = p.tw[i];
    if (nf > 0) {
        const int span = (nf - 1) * p.hop + p.n_fft;
        const int p0 = f0 * p.hop - M;                    // padded-clip coordinate of s_samples[0]
        const int jlo = max(0, -p0);                      // first j with clip index >= 0
        const int jhi = min(span, p.clip_len - p0);       // first j with clip index >= clip_len
        // reflected head / tail (only the first / last CTA of a window has any)
        for (int j = tid; j < jlo; j += kLogmelThreads) {
            long long a = start + static_cast<long long>(-(p0 + j));
            s_samples[j] = (a >= lo && a < hi) ? __ldg(p.audio + a) : 0.0f;
        }
        for (int j = jhi + tid; j < span; j += kLogmelThreads) {
            long long a = start + (2LL * (p.clip_len - 1) - (p0 + j));
            s_samples[j] = (a >= lo && a < hi) ? __ldg(p.audio + a) : 0.0f;
        }
        // interior: audio index a = abase + j, contiguous -> aligned float4 loads
        const long long abase = start + p0;
        const long long a_first = abase + jlo, a_last = abase + jhi;            // [a_first, a_last)
        const long long v_lo = max(a_first, lo), v_hi = min(a_last, hi);        // fully-valid range
        long long a4 = (a_first >= 0 ? (a_first & ~3LL) : -((-a_first + 3) & ~3LL)) + 4LL * tid;
        for (; a4 < a_last; a4 += 4LL * kLogmelThreads) {
            if (a4 >= v_lo && a4 + 4 <= v_hi) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.audio + a4));
                const int j = static_cast<int>(a4 - abase);
                s_samples[j] = v.x;
                s_samples[j + 1] = v.y;
                s_samples[j + 2] = v.z;
                s_samples[j + 3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const long long a = a4 + e;
                    if (a >= a_first && a < a_last)
                        s_samples[static_cast<int>(a - abase)] = (a >= lo && a < hi) ? __ldg(p.audio + a) : 0.0f;
                }
            }
        }
    }
    __syncthreads();

    // ---- per-frame FFT + mel + log ---------------------------------------------------------------
    const int G = p.group_threads;
    const int g = tid / G;
    const int gt = tid - g * G;
    float2* bufA = s_fft + static_cast<size_t>(g) * 2 * M;
    float2* bufB = bufA + M;
    float vmax = -INFINITY, vmin = INFINITY;
    const int iters = (nf + p.n_groups - 1) / p.n_groups;

    for (int it = 0; it < iters; ++it) {
        const int fl = it * p.n_groups + g;              // frame index local to this CTA
        const bool active = fl < nf;
        // window + pack: z[j] = x[2j] w[2j] + i x[2j+1] w[2j+1]
        if (active) {
            const float* x = s_samples + fl * p.hop;
            for (int j = gt; j < M; j += G) {
                const float2 h = __ldg(reinterpret_cast<const float2*>(p.hann) + j);
                bufA[j] = make_float2(x[2 * j] * h.x, x[2 * j + 1] * h.y);
            }
        }
        group_sync(g, G);
        float2* src = bufA;
        float2* dst = bufB;
        int ns = 1, lg = 0;
        if (p.log2m & 1) {                               // one radix-2 pass first when log2(M) is odd
            if (active) {
                const int half = M >> 1;
                for (int j = gt; j < half; j += G) {
                    const float2 u0 = src[j], u1 = src[j + half];
                    dst[2 * j] = make_float2(u0.x + u1.x, u0.y + u1.y);
                    dst[2 * j + 1] = make_float2(u0.x - u1.x, u0.y - u1.y);
                }
            }
            group_sync(g, G);
            float2* t = src; src = dst; dst = t;
            ns = 2; lg = 1;
        }
        for (; lg < p.log2m; lg += 2, ns <<= 2) {        // radix-4 Stockham passes
            if (active) {
                const int quarter = M >> 2;
                const int tw_stride = p.n_fft / (ns * 4);
                for (int j = gt; j < quarter; j += G) {
                    const int k = j & (ns - 1);
                    float2 v0 = src[j];
                    float2 v1 = src[j + quarter];
                    float2 v2 = src[j + 2 * quarter];
                    float2 v3 = src[j + 3 * quarter];
                    if (k != 0) {
                        v1 = cmul(v1, twiddle(s_tw, k * tw_stride, M));
                        v2 = cmul(v2, twiddle(s_tw, 2 * k * tw_stride, M));
                        v3 = cmul(v3, twiddle(s_tw, 3 * k * tw_stride, M));
                    }
                    const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
                    const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
                    const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
                    const float2 a3 = make_float2(v1.y - v3.y, v3.x - v1.x);       // (v1 - v3) * (-i)
                    const int j0 = ((j - k) << 2) + k;
                    dst[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
                    dst[j0 + ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
                    dst[j0 + 2 * ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
                    dst[j0 + 3 * ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
                }
            }
            group_sync(g, G);
            float2* t = src; src = dst; dst = t;
        }
        // real-FFT untangle + power spectrum: P[k], k = 0..M, into dst (as floats)
        float* power = reinterpret_cast<float*>(dst);
        if (active) {
            for (int k = gt; k <= M; k += G) {
                const float2 zk = src[k & (M - 1)];
                const float2 zm = src[(M - k) & (M - 1)];
                const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                const float2 o = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));   // (zk - conj zm)/(2i)
                const float2 t = cmul(o, twiddle(s_tw, k, M));
                const float re = e.x + t.x, im = e.y + t.y;
                power[k] = re * re + im * im;
            }
        }
        group_sync(g, G);
        // sparse mel: 8 lanes cooperate on one filter
        if (active) {
            const int sub = gt >> 3, sl = gt & 7, nsub = G >> 3;
            for (int m0 = 0; m0 < kMels; m0 += nsub) {
                const int m = m0 + sub;
                float acc = 0.0f;
                if (m < kMels) {
                    const int b0 = __ldg(p.mel_start + m), cnt = __ldg(p.mel_cnt + m);
                    const float* wt = p.mel_w + __ldg(p.mel_off + m);
                    for (int i = sl; i < cnt; i += 8) acc = fmaf(__ldg(wt + i), power[b0 + i], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                if (m < kMels && sl == 0) {
                    const float v = log10f(fmaxf(acc, 1e-10f));
                    s_tile[m * p.tile_stride + fl] = v;
                    vmax = fmaxf(vmax, v);
                    if (f0 + fl < p.n_cols) vmin = fminf(vmin, v);
                }
            }
        }
        group_sync(g, G);
    }

    // ---- window-global max / min across the cluster ---------------------------------------------
    vmax = warp_max(vmax);
    vmin = warp_min(vmin);
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = vmax;
        s_red[1][tid >> 5] = vmin;
    }
    __syncthreads();
    if (tid == 0) {
        float a = s_red[0][0], b = s_red[1][0];
        for (int i = 1; i < kLogmelThreads / 32; ++i) {
            a = fmaxf(a, s_red[0][i]);
            b = fminf(b, s_red[1][i]);
        }
        s_cluster_red[0] = a;
        s_cluster_red[1] = b;
    }
    cluster.sync();
    float gmax = -INFINITY, gmin = INFINITY;
    for (int r = 0; r < csize; ++r) {
        const float* peer = cluster.map_shared_rank(s_cluster_red, r);
        gmax = fmaxf(gmax, peer[0]);
        gmin = fminf(gmin, peer[1]);
    }
    cluster.sync();                                     // peers may exit only after everyone has read

    // ---- clamp + affine + store --------------------------------------------------------------------
    const float floor_v = gmax - 8.0f;
    float* outw = p.out + static_cast<size_t>(w) * kMels * p.n_cols;
    const int c1 = min(f1, p.n_cols);
    const int ncol = max(0, c1 - f0);
    for (int idx = tid; idx < kMels * ncol; idx += kLogmelThreads) {
        const int m = idx / ncol, c = idx - m * ncol;
        const float v = fmaxf(s_tile[m * p.tile_stride + c], floor_v);
        outw[m * p.n_cols + f0 + c] = (v + 4.0f) * 0.25f;
    }
    // clips with fewer than n_cols frames: pad with the window minimum (model.py:155-161)
    if (p.n_frames < p.n_cols) {
        const float padv = (p.n_frames > 0) ? (fmaxf(gmin, floor_v) + 4.0f) * 0.25f : 0.0f;
        const int npad = p.n_cols - p.n_frames;
        const int per = (npad + csize - 1) / csize;
        const int q0 = p.n_frames + rank * per, q1 = min(p.n_cols, q0 + per);
        const int nq = max(0, q1 - q0);
        for (int idx = tid; idx < kMels * nq; idx += kLogmelThreads) {
            const int m = idx / nq, c = idx - m * nq;
            outw[m * p.n_cols + q0 + c] = padv;
        }
    }
}

// ------------------------------------------------------------------------------------------- host
static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

int logmel_plan_create(int n_fft, int hop, int clip_len, int n_cols, const float* mel_filters_host, int n_freq,
                       LogmelPlan** out) {
    WSB_REQUIRE(n_fft >= 64 && (n_fft & (n_fft - 1)) == 0 && n_fft <= 8192, "n_fft must be a power of two in [64, 8192]");
    WSB_REQUIRE(n_freq == n_fft / 2 + 1, "mel filter bank must have n_fft/2+1 rows");
    WSB_REQUIRE(hop >= 1 && clip_len > n_fft / 2 && n_cols >= 1, "bad hop / clip_len / n_cols");
    LogmelPlan* pl = new LogmelPlan();
    pl->n_fft = n_fft;
    pl->hop = hop;
    pl->clip_len = clip_len;
    pl->n_cols = n_cols;
    pl->n_frames = clip_len / hop;
    const int M = n_fft / 2;
    pl->log2m = ilog2(M);

    // pick the largest cluster (<= 8) and the frame slice so that everything fits in shared memory
    int dev = 0, max_smem = 0;
    WSB_CHECK_CUDA(cudaGetDevice(&dev));
    WSB_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    pl->cluster = kMaxCluster;
    pl->frames_per_cta = std::max(1, ceil_div(std::max(pl->n_frames, 1), pl->cluster));
    pl->span_floats = ((pl->frames_per_cta - 1) * hop + n_fft + 3) & ~3;
    pl->tile_stride = pl->frames_per_cta | 1;
    // as many independent frame groups as shared memory allows (a warp per frame when possible): the
    // FFT passes of different frames then overlap instead of serialising on block-wide barriers
    const size_t fixed_bytes = sizeof(float) * (static_cast<size_t>(pl->span_floats) + kMels * pl->tile_stride) +
                               sizeof(float2) * static_cast<size_t>(M) + 2048;
    pl->group_threads = 32;
    while (pl->group_threads < kLogmelThreads &&
           fixed_bytes + sizeof(float2) * 2 * M * (kLogmelThreads / pl->group_threads) > static_cast<size_t>(max_smem))
        pl->group_threads *= 2;
    pl->n_groups = kLogmelThreads / pl->group_threads;
    pl->smem_bytes = sizeof(float) * (static_cast<size_t>(pl->span_floats) + kMels * pl->tile_stride) +
                     sizeof(float2) * (static_cast<size_t>(M) + static_cast<size_t>(pl->n_groups) * 2 * M);
    if (pl->smem_bytes + 1024 > static_cast<size_t>(max_smem)) {
        set_last_error("log-mel plan needs " + std::to_string(pl->smem_bytes) + " B shared memory per CTA (limit " +
                       std::to_string(max_smem) + "): hop/n_fft combination too large for an 8-CTA cluster");
        delete pl;
        return 3;
    }

    std::vector<float> hann(n_fft);
    std::vector<float2> tw(M);
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < n_fft; ++i) hann[i] = static_cast<float>(0.5 - 0.5 * cos(two_pi * i / n_fft));
    for (int i = 0; i < M; ++i)
        tw[i] = make_float2(static_cast<float>(cos(two_pi * i / n_fft)), static_cast<float>(-sin(two_pi * i / n_fft)));
    std::vector<int> st(kMels), cnt(kMels), off(kMels);
    std::vector<float> wts;
    for (int m = 0; m < kMels; ++m) {
        int a = -1, b = -1;
        for (int k = 0; k < n_freq; ++k)
            if (mel_filters_host[static_cast<size_t>(k) * kMels + m] != 0.0f) {
                if (a < 0) a = k;
                b = k;
            }
        st[m] = a < 0 ? 0 : a;
        cnt[m] = a < 0 ? 0 : b - a + 1;
        off[m] = static_cast<int>(wts.size());
        for (int k = 0; k < cnt[m]; ++k) wts.push_back(mel_filters_host[static_cast<size_t>(st[m] + k) * kMels + m]);
    }
    pl->nnz = static_cast<int>(wts.size());
    if (wts.empty()) wts.push_back(0.0f);
    WSB_CHECK_CUDA(cudaMalloc(&pl->hann, sizeof(float) * n_fft));
    WSB_CHECK_CUDA(cudaMalloc(&pl->tw, sizeof(float2) * M));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_start, sizeof(int) * kMels));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_cnt, sizeof(int) * kMels));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_off, sizeof(int) * kMels));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_w, sizeof(float) * wts.size()));
    WSB_CHECK_CUDA(cudaMemcpy(pl->hann, hann.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->tw, tw.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_start, st.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_cnt, cnt.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_off, off.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_w, wts.data(), sizeof(float) * wts.size(), cudaMemcpyHostToDevice));
    // plans with different shapes share the kernel: always opt in to the device maximum
    cudaFuncAttributes fa;
    WSB_CHECK_CUDA(cudaFuncGetAttributes(&fa, logmel_kernel));
    WSB_CHECK_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        max_smem - static_cast<int>(fa.sharedSizeBytes)));
    *out = pl;
    return 0;
}

void logmel_plan_destroy(LogmelPlan* pl) {
    if (!pl) return;
    cudaFree(pl->hann);
    cudaFree(pl->tw);
    cudaFree(pl->mel_start);
    cudaFree(pl->mel_cnt);
    cudaFree(pl->mel_off);
    cudaFree(pl->mel_w);
    delete pl;
}

int logmel_run(const LogmelPlan* pl, const float* audio_dev, const long long* win_dev, int n_win, float* out_dev,
               cudaStream_t stream) {
    if (n_win <= 0) return 0;
    WSB_REQUIRE((reinterpret_cast<uintptr_t>(audio_dev) & 15) == 0, "audio buffer must be 16-byte aligned");
    LogmelParams p;
    p.audio = audio_dev;
    p.win = win_dev;
    p.out = out_dev;
    p.hann = pl->hann;
    p.tw = pl->tw;
    p.mel_start = pl->mel_start;
    p.mel_cnt = pl->mel_cnt;
    p.mel_off = pl->mel_off;
    p.mel_w = pl->mel_w;
    p.n_fft = pl->n_fft;
    p.hop = pl->hop;
    p.clip_len = pl->clip_len;
    p.n_frames = pl->n_frames;
    p.n_cols = pl->n_cols;
    p.log2m = pl->log2m;
    p.frames_per_cta = pl->frames_per_cta;
    p.group_threads = pl->group_threads;
    p.n_groups = pl->n_groups;
    p.span_floats = pl->span_floats;
    p.tile_stride = pl->tile_stride;

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(n_win) * pl->cluster);
    cfg.blockDim = dim3(kLogmelThreads);
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, logmel_kernel, p));
    count_launch();
    return 0;
}

size_t logmel_plan_smem(const LogmelPlan* pl) { return pl->smem_bytes; }
int logmel_plan_cluster(const LogmelPlan* pl) { return pl->cluster; }

}  // namespace wsb
