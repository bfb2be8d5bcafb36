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
#include "logmel_fft.cuh"
#include "wsb_internal.h"

#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace cg = cooperative_groups;

namespace wsb {

constexpr int kLogmelThreads = 512;
constexpr int kMaxBfly = 8;          // radix-4 butterflies a thread may hold per pass
constexpr int kMels = 80;
constexpr int kMaxCluster = 8;

struct LogmelPlan {
    int n_fft, hop, clip_len, n_frames, n_cols, log2m;
    int cluster, frames_per_cta, group_threads, n_groups;
    int span_floats, tile_stride;
    int sub_frames = 0;             // one-frame-per-warp kernel: frames staged at a time (= frames_per_cta unless the span is too large)
    size_t smem_bytes;
    bool smem_tables = false;
    bool two_pass = false;          // n_fft 512 / 1024: logmel2_kernel (two-pass register FFT, lane = frame mel stage)
    struct TwoPassCfg {             // one per cluster size that fits in shared memory; logmel_run picks per launch
        int cluster, frames_per_cta, tile_stride, resident_clusters;
        size_t smem_bytes;
    };
    std::vector<TwoPassCfg> tp;
    int chunk_floats = 0;           // two-pass kernel: floats per staged sample chunk
    int chunk_bufs = 2;             // chunk buffers: 2 (the next chunk lands while this one is transformed), 1 for large hops
    int4* mel_desc = nullptr;       // [80] {first bin, groups of 4 taps, offset into mel_w4, 0} (two-pass kernel)
    float* mel_w4 = nullptr;        // CSR weights, every filter zero-padded to a multiple of 4 taps
    int nnz4 = 0;
    float* hann_half = nullptr;     // [n_fft]  0.5 * hann (two-pass kernel)
    float2* twp = nullptr;          // [16][M/16]  exp(-2 pi i n1 k2 / M) (two-pass kernel)
    float* hann = nullptr;          // [n_fft]
    float2* tw = nullptr;           // [3 n_fft/4]  exp(-2 pi i j / n_fft)
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
    int frames_per_cta, group_threads, n_groups, span_floats, tile_stride, sub_frames;
    int mel_nnz;
    const float* hann_half;         // two-pass kernel only
    const float2* twp;
    const int4* mel_desc;
    const float* mel_w4;
    int mel_nnz4;
    int chunk_floats;
    int chunk_bufs;
};

// padded index into a frame group's FFT buffer: the strided scatters of the early Stockham passes (stride 4,
// 8, 16 complex elements) would otherwise hit the same banks 8-way
__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// ---- register-resident FFT for M = 256 / 512 complex points per frame (n_fft 512 / 1024: every configuration up to 80 kHz)
// Cooley-Tukey M = 32 x R: lane n1 holds the R points x[n1 + 32 n2]; an R-point DIF in registers, one twiddle
// W_M^(n1 k2), a 32-point DIF ACROSS the lanes with shuffles (5 stages), and the spectrum is written once, transposed and
// padded, for the real-FFT untangle.  Two shared-memory exchanges per frame instead of the Stockham path's five passes
// with two barriers each (round-1 ncu: 179 M shared-memory bank conflicts per launch, IPC 0.9).
__device__ constexpr float kCos16[16] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f,
                                         -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.0f,
                                         -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f, 0.0f,
                                         0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
__device__ constexpr float kSin16[16] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f, 1.0f,
                                         0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f,
                                         -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.0f,
                                         -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
__host__ __device__ constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
// in-register R-point DIF (radix 2): v[i] <- X[bitrev(i)]
template <int R>
__device__ __forceinline__ void fft_dif_registers(float2 (&v)[R]) {
#pragma unroll
    for (int h = R / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & h) == 0) {
                const float2 a = v[i], b = v[i + h];
                v[i] = make_float2(a.x + b.x, a.y + b.y);
                const float2 dlt = make_float2(a.x - b.x, a.y - b.y);
                const int k16 = (i & (h - 1)) * (8 / h);            // exp(-2 pi i t / (2h)) = (cos, -sin)(2 pi k16 / 16)
                if (k16 == 0) v[i + h] = dlt;
                else if (k16 == 4) v[i + h] = make_float2(dlt.y, -dlt.x);
                else v[i + h] = make_float2(dlt.x * kCos16[k16] + dlt.y * kSin16[k16], dlt.y * kCos16[k16] - dlt.x * kSin16[k16]);
            }
        }
    }
}

// Compile-time shape of the per-frame FFT: M = n_fft/2 complex points handled by a group of G threads
// (a warp up to n_fft 1024; larger groups beyond), NB radix-4 butterflies per thread and pass.
template <int LOG2M>
struct FftCfg {
    static constexpr int M = 1 << LOG2M;
    static constexpr int G = LOG2M <= 9 ? 32 : LOG2M == 10 ? 64 : LOG2M == 11 ? 256 : 512;   // large FFTs: fewer, wider groups (smem)
    static constexpr int NB = M / (4 * G);
    static constexpr int NGROUPS = kLogmelThreads / G;
    static constexpr int MPAD = M + M / 16;                 // FFT buffer with one pad element per 16 (bank conflicts)
    static constexpr int GROUP_FLOATS = 2 * MPAD + M + 4;   // padded complex points + (M+4) floats of power spectrum
    static constexpr int TW = 3 * M / 2;                    // twiddle table entries: exp(-2 pi i j / n_fft), j < 3 n_fft / 4
};

// barrier among the G threads that cooperate on one frame (a warp, or a named barrier per group), so
// the frame groups of a CTA run their FFT passes independently of each other
template <int G>
__device__ __forceinline__ void group_sync(int g) {
    if constexpr (G == 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(G) : "memory");
    }
}

// kSmemTables: Hann window, twiddles and the CSR mel bank are staged in shared memory (fits up to
// n_fft 2048 for the usual hops); otherwise they are read through L1 from global memory.
template <int LOG2M, bool kSmemTables>
__global__ void __launch_bounds__(kLogmelThreads) logmel_kernel(const LogmelParams p) {
    using C = FftCfg<LOG2M>;
    constexpr int M = C::M, G = C::G, NB = C::NB, N_FFT = 2 * C::M;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int csize = static_cast<int>(cluster.num_blocks());
    const int w = blockIdx.x / csize;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_samples = reinterpret_cast<float*>(smem_raw);                       // span_floats
    float* s_tile = s_samples + p.span_floats;                                   // 80 * tile_stride (+ pad to 4)
    float* s_fft = s_tile + ((kMels * p.tile_stride + 3) & ~3);                  // NGROUPS * GROUP_FLOATS
    float* s_tab = s_fft + C::NGROUPS * C::GROUP_FLOATS;                         // optional tables
    __shared__ float s_red[2][kLogmelThreads / 32];
    __shared__ float s_cluster_red[2];                                           // this CTA's {max, min}, read by peers
    __shared__ int s_mel[3][kMels];                                              // CSR: first bin, count, weight offset

    const float2* tw;
    const float* hann;
    const float* mel_w;
    if constexpr (kSmemTables) {
        float2* t = reinterpret_cast<float2*>(s_tab);
        float* h = s_tab + 2 * C::TW;
        float* mw = h + N_FFT;
        for (int i = tid; i < C::TW; i += kLogmelThreads) t[i] = p.tw[i];
        for (int i = tid; i < N_FFT; i += kLogmelThreads) h[i] = p.hann[i];
        for (int i = tid; i < p.mel_nnz; i += kLogmelThreads) mw[i] = p.mel_w[i];
        tw = t;
        hann = h;
        mel_w = mw;
    } else {
        tw = p.tw;
        hann = p.hann;
        mel_w = p.mel_w;
    }
    if (tid < kMels) {
        s_mel[0][tid] = p.mel_start[tid];
        s_mel[1][tid] = p.mel_cnt[tid];
        s_mel[2][tid] = p.mel_off[tid];
    }

    const int f0 = rank * p.frames_per_cta;
    const int f1 = min(p.n_frames, f0 + p.frames_per_cta);
    const int nf = max(0, f1 - f0);

    const long long start = p.win[3 * w + 0];
    const long long lo = p.win[3 * w + 1];
    const long long hi = p.win[3 * w + 2];

    // ---- per-frame FFT + mel + log ---------------------------------------------------------------
    const int g = tid / G;
    const int gt = tid - g * G;
    float2* buf = reinterpret_cast<float2*>(s_fft + g * C::GROUP_FLOATS);
    float* power = reinterpret_cast<float*>(buf + C::MPAD);
    float vmax = -INFINITY, vmin = INFINITY;
    // twiddles of the cross-lane DIF stages (register FFT path): exp(-2 pi i (lane mod h) / (2h)), h = 16, 8, 4, 2, 1
    float2 lane_tw[5];
#pragma unroll
    for (int st = 0; st < 5; ++st) {
        const int h = 16 >> st;
        float sn, cs;
        sincospif(static_cast<float>(gt & (h - 1)) / static_cast<float>(h), &sn, &cs);
        lane_tw[st] = make_float2(cs, -sn);
    }
    constexpr int Q = M / 4;

#pragma unroll 1
    for (int sf = 0; sf < nf; sf += p.sub_frames) {       // rounds of sub_frames frames (one round unless the hop is large)
    const int nfs = min(p.sub_frames, nf - sf);
    __syncthreads();                                      // the previous round is done with s_samples
    // ---- stage the sample span of this round's frames (all of the CTA's frames unless the span is too large) -------------
    if (nfs > 0) {
        const int span = (nfs - 1) * p.hop + N_FFT;
        const int p0 = (f0 + sf) * p.hop - M;             // padded-clip coordinate of s_samples[0]
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
        long long a4 = (a_first & ~3LL) + 4LL * tid;
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

    const int iters = (nfs + C::NGROUPS - 1) / C::NGROUPS;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int fls = it * C::NGROUPS + g;             // frame index local to this round
        const int fl = sf + fls;                         // frame index local to this CTA
        const bool active = fls < nfs;
        if constexpr (LOG2M == 8 || LOG2M == 9) {
            constexpr int R = M / 32, LOGR = LOG2M - 5;
            // transposed, padded spectrum: Z[R k1 + k2] lives at zt[k2 * 33 + k1]
            float2* zt = buf;
            if (active) {
                const float* x = s_samples + fls * p.hop;
                const bool x_aligned = ((fls * p.hop) & 1) == 0;       // warp-uniform
                float2 v[R];
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) {
                    const int j = gt + 32 * n2;
                    const float2 h = reinterpret_cast<const float2*>(hann)[j];
                    const float2 xv = x_aligned ? reinterpret_cast<const float2*>(x)[j] : make_float2(x[2 * j], x[2 * j + 1]);
                    v[n2] = make_float2(xv.x * h.x, xv.y * h.y);
                }
                fft_dif_registers<R>(v);
#pragma unroll
                for (int i = 1; i < R; ++i) {                          // W_M^(n1 k2), k2 = bitrev(i); tw[] holds W_(2M)^j, j < 3M/2
                    constexpr int kk = 0;
                    (void)kk;
                    int e = gt * bitrev(i, LOGR);
                    const bool neg = e >= M / 2;
                    if (neg) e -= M / 2;
                    float2 w = tw[2 * e];
                    if (neg) w = make_float2(-w.x, -w.y);
                    v[i] = cmul(v[i], w);
                }
#pragma unroll
                for (int h = 16; h >= 1; h >>= 1) {                    // 32-point DIF across the lanes
                    const bool upper = (gt & h) != 0;
                    const float2 wl = lane_tw[h == 16 ? 0 : h == 8 ? 1 : h == 4 ? 2 : h == 2 ? 3 : 4];
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const float ox = __shfl_xor_sync(0xffffffffu, v[i].x, h);
                        const float oy = __shfl_xor_sync(0xffffffffu, v[i].y, h);
                        if (upper) {
                            const float2 dlt = make_float2(ox - v[i].x, oy - v[i].y);
                            v[i] = (h == 1) ? dlt : cmul(dlt, wl);
                        } else {
                            v[i] = make_float2(v[i].x + ox, v[i].y + oy);
                        }
                    }
                }
                const int k1 = bitrev(gt, 5);
#pragma unroll
                for (int i = 0; i < R; ++i) zt[bitrev(i, LOGR) * 33 + k1] = v[i];
            }
            group_sync<G>(g);
            if (active) {                                              // real-FFT untangle + power spectrum: P[k], k = 0..M
#pragma unroll
                for (int i = 0; i <= M / G; ++i) {
                    const int k = gt + i * G;
                    if (i < M / G || k == M) {
                        const int ka = k & (M - 1), kb = (M - k) & (M - 1);
                        const float2 zk = zt[(ka & (R - 1)) * 33 + (ka >> LOGR)];
                        const float2 zm = zt[(kb & (R - 1)) * 33 + (kb >> LOGR)];
                        const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                        const float2 o = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));   // (zk - conj zm)/(2i)
                        const float2 t = cmul(o, tw[k]);
                        const float re = e.x + t.x, im = e.y + t.y;
                        power[k] = re * re + im * im;
                    }
                }
            }
            group_sync<G>(g);
        } else {
            // window + pack: z[j] = x[2j] w[2j] + i x[2j+1] w[2j+1]
            if (active) {
                const float* x = s_samples + fls * p.hop;
                const bool x_aligned = ((fls * p.hop) & 1) == 0;       // warp-uniform
    #pragma unroll
                for (int i = 0; i < M / G; ++i) {
                    const int j = gt + i * G;
                    const float2 h = reinterpret_cast<const float2*>(hann)[j];
                    float2 xv;
                    if (x_aligned) {
                        xv = reinterpret_cast<const float2*>(x)[j];
                    } else {
                        xv = make_float2(x[2 * j], x[2 * j + 1]);
                    }
                    buf[fpad(j)] = make_float2(xv.x * h.x, xv.y * h.y);
                }
            }
            group_sync<G>(g);
            // In-place Stockham passes: every thread pulls the inputs of all its butterflies into registers,
            // the group synchronises, then the outputs are scattered -- one M-point buffer per frame group.
            if constexpr (LOG2M & 1) {                       // one radix-2 pass first when log2(M) is odd
                float2 u[2 * NB][2];
                if (active) {
    #pragma unroll
                    for (int i = 0; i < 2 * NB; ++i) {
                        const int j = gt + i * G;
                        u[i][0] = buf[fpad(j)];
                        u[i][1] = buf[fpad(j + M / 2)];
                    }
                }
                group_sync<G>(g);
                if (active) {
    #pragma unroll
                    for (int i = 0; i < 2 * NB; ++i) {
                        const int j = gt + i * G;
                        buf[fpad(2 * j)] = make_float2(u[i][0].x + u[i][1].x, u[i][0].y + u[i][1].y);
                        buf[fpad(2 * j + 1)] = make_float2(u[i][0].x - u[i][1].x, u[i][0].y - u[i][1].y);
                    }
                }
                group_sync<G>(g);
            }
    #pragma unroll
            for (int lg = (LOG2M & 1); lg < LOG2M; lg += 2) {      // radix-4 passes, ns = 2^lg
                const int ns = 1 << lg;
                const int tw_stride = N_FFT / (ns * 4);
                float2 v[NB][4];
                if (active) {
    #pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        const int j = gt + i * G;
                        v[i][0] = buf[fpad(j)];
                        v[i][1] = buf[fpad(j + Q)];
                        v[i][2] = buf[fpad(j + 2 * Q)];
                        v[i][3] = buf[fpad(j + 3 * Q)];
                    }
                }
                group_sync<G>(g);
                if (active) {
    #pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        const int j = gt + i * G;
                        const int k = j & (ns - 1);
                        float2 v1 = v[i][1], v2 = v[i][2], v3 = v[i][3];
                        if (lg > 0) {                              // ns == 1: all twiddles are 1
                            v1 = cmul(v1, tw[k * tw_stride]);
                            v2 = cmul(v2, tw[2 * k * tw_stride]);
                            v3 = cmul(v3, tw[3 * k * tw_stride]);
                        }
                        const float2 v0 = v[i][0];
                        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
                        const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
                        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
                        const float2 a3 = make_float2(v1.y - v3.y, v3.x - v1.x);       // (v1 - v3) * (-i)
                        const int j0 = ((j - k) << 2) + k;
                        buf[fpad(j0)] = make_float2(a0.x + a2.x, a0.y + a2.y);
                        buf[fpad(j0 + ns)] = make_float2(a1.x + a3.x, a1.y + a3.y);
                        buf[fpad(j0 + 2 * ns)] = make_float2(a0.x - a2.x, a0.y - a2.y);
                        buf[fpad(j0 + 3 * ns)] = make_float2(a1.x - a3.x, a1.y - a3.y);
                    }
                }
                group_sync<G>(g);
            }
            // real-FFT untangle + power spectrum: P[k], k = 0..M
            if (active) {
    #pragma unroll
                for (int i = 0; i <= M / G; ++i) {
                    const int k = gt + i * G;
                    if (i < M / G || k == M) {
                        const float2 zk = buf[fpad(k & (M - 1))];
                        const float2 zm = buf[fpad((M - k) & (M - 1))];
                        const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                        const float2 o = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));   // (zk - conj zm)/(2i)
                        const float2 t = cmul(o, tw[k]);
                        const float re = e.x + t.x, im = e.y + t.y;
                        power[k] = re * re + im * im;
                    }
                }
            }
            group_sync<G>(g);
        }
        // sparse mel: each thread owns filters gt, gt+G, ... (<= 2 triangles per FFT bin, CSR rows)
        if (active) {
            for (int m = gt; m < kMels; m += G) {
                const int b0 = s_mel[0][m], cnt = s_mel[1][m];
                const float* wt = mel_w + s_mel[2][m];
                float acc0 = 0.0f, acc1 = 0.0f;
                int i = 0;
                for (; i + 1 < cnt; i += 2) {
                    acc0 = fmaf(wt[i], power[b0 + i], acc0);
                    acc1 = fmaf(wt[i + 1], power[b0 + i + 1], acc1);
                }
                if (i < cnt) acc0 = fmaf(wt[i], power[b0 + i], acc0);
                const float v = log10f(fmaxf(acc0 + acc1, 1e-10f));
                s_tile[m * p.tile_stride + fl] = v;
                vmax = fmaxf(vmax, v);
                if (f0 + fl < p.n_cols) vmin = fminf(vmin, v);
            }
        }
        group_sync<G>(g);
    }
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

// ---------------------------------------------------------------------------------------------------------------------
// Two-pass kernel for n_fft 512 / 1024 (every sampling rate up to 80 kHz).  Same contract and the same cluster-per-window
// layout as logmel_kernel above; what changes is how a CTA spends its instructions (round-2 ncu of the one-frame-per-warp
// kernel: ~2600 warp instructions per frame, 960 of them in the five shuffle stages, 600 in the untangle, 400 in the mel
// stage):
//   * FFT: the two register passes of logmel_fft.cuh with one shared-memory transpose between them, no shuffles
//   * untangle: one thread per (k, M - k) pair, powers written transposed, P[k][frame]
//   * mel + log10 with lane = frame over the 16 (M = 512) or 32 (M = 256) frames of a CTA iteration: every lane of a warp
//     walks the same filter, the weights are warp-uniform loads and the power reads are conflict-free
//   * samples are staged per iteration (two cp.async buffers) instead of per CTA, which leaves room for 4-CTA clusters
//     (8-CTA clusters keep 15 x 8 = 120 of the 148 SMs busy; the plan picks the cluster size per launch from
//     cudaOccupancyMaxActiveClusters)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Stages the samples of window frames [frame0, frame0 + nfr) into dst (asynchronously); the first frame's first sample lands
// at dst[0].  Interior, fully valid chunks whose first sample is 16-byte aligned in the audio buffer (always the case when
// hop and the window length are multiples of 4 samples) are copied as 16-byte pieces; chunks that touch the reflected head /
// tail of the clip or the zero padding outside [lo, hi), and unaligned ones, go element by element.
template <int N_FFT>
__device__ __forceinline__ void stage_chunk(const LogmelParams& p, float* dst, int frame0, int nfr, long long start, long long lo,
                                            long long hi, int tid) {
    const int span = (nfr - 1) * p.hop + N_FFT;
    const int p0 = frame0 * p.hop - N_FFT / 2;              // padded-clip coordinate of the chunk's first sample
    const long long a0 = start + p0;
    const int n4 = (span + 3) >> 2;
    if (p0 >= 0 && p0 + span <= p.clip_len && (a0 & 3) == 0 && a0 >= lo && a0 + 4LL * n4 <= hi) {
        for (int q = tid; q < n4; q += kLogmelThreads) cp_async16(dst + 4 * q, p.audio + a0 + 4 * q);
        return;
    }
    for (int j = tid; j < span; j += kLogmelThreads) {
        int pc = p0 + j;
        if (pc < 0) pc = -pc;
        else if (pc >= p.clip_len) pc = 2 * (p.clip_len - 1) - pc;
        const long long a = start + pc;
        if (a >= lo && a < hi) cp_async4(dst + j, p.audio + a);
        else dst[j] = 0.0f;
    }
}

template <int LOG2M, int BUFS>
__global__ void __launch_bounds__(kLogmelThreads, 1) logmel2_kernel(const LogmelParams p) {
    using T = lfft::TwoPass<LOG2M>;
    constexpr int M = T::M, N_FFT = 2 * M, G = T::G;
    constexpr int NF = 16 * G;                            // frames per CTA iteration
    constexpr int PS = NF + 1;                            // row stride of the transposed power array
    constexpr int NPH = kLogmelThreads / NF;              // mel stage: filter phases (thread = (frame, phase))
    constexpr int NWARPS = kLogmelThreads / 32;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int csize = static_cast<int>(cluster.num_blocks());
    const int w = blockIdx.x / csize;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_samp = reinterpret_cast<float*>(smem_raw);                                   // chunk_bufs x chunk_floats
    float* s_tile = s_samp + BUFS * p.chunk_floats;                                          // 80 x tile_stride
    float2* s_buf = reinterpret_cast<float2*>(s_tile + ((kMels * p.tile_stride + 3) & ~3));   // per warp: T::BUF
    float* s_pt = reinterpret_cast<float*>(s_buf + NWARPS * T::BUF);                      // (M + 4) x PS, rows > M stay zero
    float2* s_tw = reinterpret_cast<float2*>(s_pt + (((M + 4) * PS + 3) & ~3));           // untangle twiddles
    float2* s_twp = s_tw + ((T::TWN + 1) & ~1);                                           // pass-1 twiddles [k2][n1]
    float* s_hann = reinterpret_cast<float*>(s_twp + T::TWP);                             // 0.5 x Hann
    float* s_melw = s_hann + N_FFT;                                                       // CSR mel weights, 4-tap groups
    __shared__ float s_red[2][NWARPS];
    __shared__ float s_cluster_red[2];
    __shared__ __align__(16) int4 s_meld[kMels];                                          // {first bin, 4-tap groups, offset}

    for (int i = tid; i < T::TWN; i += kLogmelThreads) s_tw[i] = p.tw[i];
    for (int i = tid; i < T::TWP; i += kLogmelThreads) s_twp[i] = p.twp[i];
    for (int i = tid; i < N_FFT; i += kLogmelThreads) s_hann[i] = p.hann_half[i];
    for (int i = tid; i < p.mel_nnz4; i += kLogmelThreads) s_melw[i] = p.mel_w4[i];
    for (int i = tid; i < 3 * PS; i += kLogmelThreads) s_pt[(M + 1) * PS + i] = 0.0f;     // read by the zero-weight pad taps
    if (tid < kMels) s_meld[tid] = p.mel_desc[tid];

    const int f0 = rank * p.frames_per_cta;
    const int f1 = min(p.n_frames, f0 + p.frames_per_cta);
    const int nf = max(0, f1 - f0);
    const long long start = p.win[3 * w + 0];
    const long long lo = p.win[3 * w + 1];
    const long long hi = p.win[3 * w + 2];

    const int iters = (nf + NF - 1) / NF;
    if (iters > 0) stage_chunk<N_FFT>(p, s_samp, f0, min(NF, nf), start, lo, hi, tid);
    cp_async_commit();
    float2* buf = s_buf + warp * T::BUF;
    const int slot0 = warp * G;                           // this warp's first frame slot of an iteration
    const int slot = slot0 + T::frame_of(lane);
    float vmax = -INFINITY, vmin = INFINITY;
    __syncthreads();                                      // tables staged
    float2 hwin[16];                                      // this lane's window pairs, register-resident across the frames
    T::load_window(lane, s_hann, hwin);

#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const float* cur = s_samp;
        if constexpr (BUFS == 2) {                       // the next chunk lands in the other buffer while this one is transformed
            cur += (it & 1) * p.chunk_floats;
            if (it + 1 < iters)
                stage_chunk<N_FFT>(p, s_samp + ((it + 1) & 1) * p.chunk_floats, f0 + (it + 1) * NF, min(NF, nf - (it + 1) * NF),
                                   start, lo, hi, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {                                         // large hop: one buffer, staged here (every pass-1 read of the previous
            if (it > 0)                                  // chunk happened before the barrier that ended its untangle)
                stage_chunk<N_FFT>(p, s_samp, f0 + it * NF, min(NF, nf - it * NF), start, lo, hi, tid);
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();          // chunk `it` (first pass: the tables too) is visible; the previous mel stage is done with s_pt
        const bool active = it * NF + slot < nf;
        if (active) {
            const int xo = slot * p.hop;                 // even hop (kernel-uniform): every frame is float2-aligned
            if ((p.hop & 1) == 0) T::pass1(lane, cur + xo, true, hwin, s_twp, buf);
            else T::pass1(lane, cur + xo, (xo & 1) == 0, hwin, s_twp, buf);
        }
        __syncwarp();
        float2 u[16];
        if (active) T::pass2_load(lane, buf, u);
        __syncwarp();
        if (active) T::pass2_store(lane, u, buf);
        __syncwarp();
        T::untangle(lane, buf, s_tw, s_pt, PS, slot0, active);
        __syncthreads();
        {   // mel + log10, lane = frame
            const int f = tid & (NF - 1);
            const int fl = it * NF + f;
            if (fl < nf) {
                const float* pcol = s_pt + f;
                // filters are dealt to the NPH phases widest first, alternating direction every round, so that the phases
                // (warps) carry about the same number of filter taps
                const int phase = tid / NF;
                for (int j = 0;; ++j) {
                    const int r = j * NPH + ((j & 1) ? NPH - 1 - phase : phase);
                    if (r >= kMels) break;
                    const int m = kMels - 1 - r;
                    const int4 dsc = s_meld[m];
                    const float4* wt = reinterpret_cast<const float4*>(s_melw + dsc.z);
                    const float* pw = pcol + dsc.x * PS;
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
                    for (int g4 = 0; g4 < dsc.y; ++g4, pw += 4 * PS) {
                        const float4 wv = wt[g4];
                        acc0 = fmaf(wv.x, pw[0], acc0);
                        acc1 = fmaf(wv.y, pw[PS], acc1);
                        acc0 = fmaf(wv.z, pw[2 * PS], acc0);
                        acc1 = fmaf(wv.w, pw[3 * PS], acc1);
                    }
                    // log10 through lg2.approx (absolute error ~2e-7 in log2, far inside the 1e-4 bar); the 1e-10 floor is
                    // selected, not computed, so that silence stays exactly -10
                    const float acc = acc0 + acc1;
                    const float v = acc > 1e-10f ? __log2f(acc) * 0.30102999566398120f : -10.0f;
                    s_tile[m * p.tile_stride + fl] = v;
                    vmax = fmaxf(vmax, v);
                    if (f0 + fl < p.n_cols) vmin = fminf(vmin, v);
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- window-global max / min across the cluster ---------------------------------------------
    vmax = warp_max(vmax);
    vmin = warp_min(vmin);
    if (lane == 0) {
        s_red[0][warp] = vmax;
        s_red[1][warp] = vmin;
    }
    __syncthreads();
    if (tid == 0) {
        float a = s_red[0][0], b = s_red[1][0];
        for (int i = 1; i < NWARPS; ++i) {
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
    const int ncol = max(0, min(f1, p.n_cols) - f0);
    for (int m = warp; m < kMels; m += NWARPS) {
        const float* row = s_tile + m * p.tile_stride;
        float* orow = outw + m * p.n_cols + f0;
        for (int c = lane; c < ncol; c += 32) orow[c] = (fmaxf(row[c], floor_v) + 4.0f) * 0.25f;
    }
    // clips with fewer than n_cols frames: pad with the window minimum (model.py:155-161)
    if (p.n_frames < p.n_cols) {
        const float padv = (p.n_frames > 0) ? (fmaxf(gmin, floor_v) + 4.0f) * 0.25f : 0.0f;
        const int npad = p.n_cols - p.n_frames;
        const int per = (npad + csize - 1) / csize;
        const int q0 = p.n_frames + rank * per, q1 = min(p.n_cols, q0 + per);
        const int nq = max(0, q1 - q0);
        for (int m = warp; m < kMels; m += NWARPS)
            for (int c = lane; c < nq; c += 32) outw[m * p.n_cols + q0 + c] = padv;
    }
}

typedef void (*LogmelKernelFn)(const LogmelParams);
static LogmelKernelFn pick_logmel_kernel(int log2m, bool smem_tables) {
    switch (log2m) {
        case 7: return smem_tables ? logmel_kernel<7, true> : logmel_kernel<7, false>;
        case 8: return smem_tables ? logmel_kernel<8, true> : logmel_kernel<8, false>;
        case 9: return smem_tables ? logmel_kernel<9, true> : logmel_kernel<9, false>;
        case 10: return smem_tables ? logmel_kernel<10, true> : logmel_kernel<10, false>;
        case 11: return smem_tables ? logmel_kernel<11, true> : logmel_kernel<11, false>;
        case 12: return smem_tables ? logmel_kernel<12, true> : logmel_kernel<12, false>;
        default: return nullptr;
    }
}
static int fft_group_threads(int log2m) { return log2m <= 9 ? 32 : log2m == 10 ? 64 : log2m == 11 ? 256 : 512; }

// ------------------------------------------------------------------------------------------- host
static int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

static LogmelKernelFn pick_two_pass_kernel(int log2m, int bufs) {
    if (log2m == 8) return bufs == 2 ? logmel2_kernel<8, 2> : logmel2_kernel<8, 1>;
    if (log2m == 9) return bufs == 2 ? logmel2_kernel<9, 2> : logmel2_kernel<9, 1>;
    return nullptr;
}

// Two-pass kernel (n_fft 512 / 1024): tables and one launch configuration per cluster size that fits in shared memory.
// WSB_LOGMEL_V1=1 keeps the one-frame-per-warp kernel; WSB_LOGMEL_CLUSTER=n pins the cluster size (diagnostics).
static int two_pass_plan(LogmelPlan* pl, const std::vector<float>& hann, const std::vector<int>& mel_start,
                         const std::vector<int>& mel_cnt, const std::vector<int>& mel_off, const std::vector<float>& mel_w,
                         int max_smem) {
    LogmelKernelFn fn = pick_two_pass_kernel(pl->log2m, 2);
    if (fn == nullptr || std::getenv("WSB_LOGMEL_V1") != nullptr) return 0;
    // CSR filter bank in groups of 4 taps (one float4 weight load per group); pad taps carry weight 0 and may point at
    // the three zero rows behind the last bin
    std::vector<int4> desc(kMels);
    std::vector<float> w4;
    for (int m = 0; m < kMels; ++m) {
        const int groups = (mel_cnt[m] + 3) / 4;
        desc[m] = make_int4(mel_start[m], groups, static_cast<int>(w4.size()), 0);
        for (int i = 0; i < 4 * groups; ++i) w4.push_back(i < mel_cnt[m] ? mel_w[mel_off[m] + i] : 0.0f);
    }
    if (w4.empty()) w4.resize(4, 0.0f);
    pl->nnz4 = static_cast<int>(w4.size());
    const int n_fft = pl->n_fft, M = n_fft / 2, N1 = M / 16, G = 32 / N1, NF = 16 * G, PS = NF + 1;
    const int twn = M / 2 + 1, twp_n = 16 * N1;
    pl->chunk_floats = ((NF - 1) * pl->hop + n_fft + 3) & ~3;
    cudaFuncAttributes fa;
    WSB_CHECK_CUDA(cudaFuncGetAttributes(&fa, fn));
    const size_t avail = static_cast<size_t>(max_smem) - fa.sharedSizeBytes;
    WSB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(avail)));
    WSB_CHECK_CUDA(cudaFuncSetAttribute(pick_two_pass_kernel(pl->log2m, 1), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(avail)));
    int dev = 0, n_sms = 148;
    WSB_CHECK_CUDA(cudaGetDevice(&dev));
    WSB_CHECK_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    int pinned = 0;
    if (const char* e = std::getenv("WSB_LOGMEL_CLUSTER")) pinned = std::atoi(e);
    // two chunk buffers if any cluster size fits with them, else one (hops of several hundred samples)
    for (int bufs = 2; bufs >= 1 && pl->tp.empty(); --bufs) {
    pl->chunk_bufs = bufs;
    for (int c = kMaxCluster; c >= 1; --c) {
        if (pinned > 0 && c != pinned) continue;
        LogmelPlan::TwoPassCfg cfg;
        cfg.cluster = c;
        cfg.frames_per_cta = std::max(1, ceil_div(std::max(pl->n_frames, 1), c));
        cfg.tile_stride = cfg.frames_per_cta | 1;
        const size_t floats = static_cast<size_t>(pl->chunk_bufs) * pl->chunk_floats + ((kMels * cfg.tile_stride + 3) & ~3) +
                              static_cast<size_t>(kLogmelThreads / 32) * 2 * (32 * 17) + (((M + 4) * PS + 3) & ~3) +
                              2 * ((twn + 1) & ~1) + 2 * twp_n + n_fft + pl->nnz4;
        cfg.smem_bytes = sizeof(float) * floats;
        if (cfg.smem_bytes > avail) continue;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(static_cast<unsigned>(c) * n_sms);
        lc.blockDim = dim3(kLogmelThreads);
        lc.dynamicSmemBytes = cfg.smem_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = c;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
        int resident = 0;
        if (cudaOccupancyMaxActiveClusters(&resident, pick_two_pass_kernel(pl->log2m, bufs), &lc) != cudaSuccess || resident < 1) {
            (void)cudaGetLastError();
            resident = std::max(1, n_sms / (c < 3 ? c : c + 1));        // clusters do not span GPCs: assume some loss
        }
        cfg.resident_clusters = resident;
        if (std::getenv("WSB_LOGMEL_DEBUG"))
            fprintf(stderr, "[wsb] log-mel two-pass: cluster %d, %d frames per CTA, %d chunk buffer(s), %zu B shared memory, %d clusters resident\n",
                    c, cfg.frames_per_cta, pl->chunk_bufs, cfg.smem_bytes, resident);
        pl->tp.push_back(cfg);
    }
    }
    if (pl->tp.empty()) return 0;                          // very large hop: the one-frame-per-warp kernel stays
    std::vector<float> hh(n_fft);
    for (int i = 0; i < n_fft; ++i) hh[i] = 0.5f * hann[i];
    std::vector<float2> twp(twp_n);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k2 = 0; k2 < 16; ++k2)
        for (int n1 = 0; n1 < N1; ++n1) {
            const double a = two_pi * static_cast<double>((n1 * k2) % M) / M;
            twp[k2 * N1 + n1] = make_float2(static_cast<float>(cos(a)), static_cast<float>(-sin(a)));
        }
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_desc, sizeof(int4) * kMels));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_w4, sizeof(float) * w4.size()));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_desc, desc.data(), sizeof(int4) * kMels, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_w4, w4.data(), sizeof(float) * w4.size(), cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMalloc(&pl->hann_half, sizeof(float) * n_fft));
    WSB_CHECK_CUDA(cudaMalloc(&pl->twp, sizeof(float2) * twp_n));
    WSB_CHECK_CUDA(cudaMemcpy(pl->hann_half, hh.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->twp, twp.data(), sizeof(float2) * twp_n, cudaMemcpyHostToDevice));
    pl->two_pass = true;
    return 0;
}

int logmel_plan_create(int n_fft, int hop, int clip_len, int n_cols, const float* mel_filters_host, int n_freq,
                       LogmelPlan** out) {
    WSB_REQUIRE(n_fft >= 256 && (n_fft & (n_fft - 1)) == 0 && n_fft <= 8192, "n_fft must be a power of two in [256, 8192]");
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
    pl->sub_frames = pl->frames_per_cta;
    pl->span_floats = ((pl->sub_frames - 1) * hop + n_fft + 3) & ~3;
    pl->tile_stride = pl->frames_per_cta | 1;
    std::vector<float> hann(n_fft);
    const int n_tw = 3 * n_fft / 4;
    std::vector<float2> tw(n_tw);
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < n_fft; ++i) hann[i] = static_cast<float>(0.5 - 0.5 * cos(two_pi * i / n_fft));
    for (int i = 0; i < n_tw; ++i)
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
    WSB_CHECK_CUDA(cudaMalloc(&pl->tw, sizeof(float2) * n_tw));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_start, sizeof(int) * kMels));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_cnt, sizeof(int) * kMels));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_off, sizeof(int) * kMels));
    WSB_CHECK_CUDA(cudaMalloc(&pl->mel_w, sizeof(float) * wts.size()));
    WSB_CHECK_CUDA(cudaMemcpy(pl->hann, hann.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->tw, tw.data(), sizeof(float2) * n_tw, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_start, st.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_cnt, cnt.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_off, off.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice));
    WSB_CHECK_CUDA(cudaMemcpy(pl->mel_w, wts.data(), sizeof(float) * wts.size(), cudaMemcpyHostToDevice));
    // shared-memory budget: samples + tile + per-group FFT buffers (+ tables when they fit)
    pl->group_threads = fft_group_threads(pl->log2m);
    pl->n_groups = kLogmelThreads / pl->group_threads;
    const size_t static_bytes = 2048;                   // s_red, s_cluster_red, s_mel + margin
    const size_t fixed_bytes = sizeof(float) * (((kMels * pl->tile_stride + 3) & ~3) +
                                                static_cast<size_t>(pl->n_groups) * (2 * (M + M / 16) + M + 4));
    // large hops: the CTA's frames are staged in rounds of sub_frames frames (whole iterations of n_groups frames) until the
    // span fits -- e.g. 192 kHz / 0.0025 s (hop 480, n_fft 4096) runs 2 rounds
    while (sizeof(float) * pl->span_floats + fixed_bytes + static_bytes > static_cast<size_t>(max_smem) && pl->sub_frames > pl->n_groups) {
        pl->sub_frames = std::max(pl->n_groups, ceil_div(ceil_div(pl->sub_frames, 2), pl->n_groups) * pl->n_groups);
        pl->span_floats = ((pl->sub_frames - 1) * hop + n_fft + 3) & ~3;
    }
    const size_t base_bytes = sizeof(float) * pl->span_floats + fixed_bytes;
    const size_t table_bytes = sizeof(float) * (2 * static_cast<size_t>(n_tw) + n_fft + ((wts.size() + 3) & ~size_t(3)));
    pl->smem_tables = base_bytes + table_bytes + static_bytes <= static_cast<size_t>(max_smem);
    pl->smem_bytes = base_bytes + (pl->smem_tables ? table_bytes : 0);
    // the one-frame-per-warp kernel stages a CTA's whole sample span: hops of a few hundred samples do not fit; the two-pass
    // kernel (n_fft 512 / 1024) stages per iteration and reaches further
    const bool v1_fits = pl->smem_bytes + static_bytes <= static_cast<size_t>(max_smem);
    LogmelKernelFn fn = pick_logmel_kernel(pl->log2m, pl->smem_tables);
    WSB_REQUIRE(fn != nullptr, "unsupported n_fft");
    if (v1_fits) {
        cudaFuncAttributes fa;
        WSB_CHECK_CUDA(cudaFuncGetAttributes(&fa, fn));
        WSB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            max_smem - static_cast<int>(fa.sharedSizeBytes)));
    }
    if (two_pass_plan(pl, hann, st, cnt, off, wts, max_smem)) {
        logmel_plan_destroy(pl);
        return 1;
    }
    if (!v1_fits && !pl->two_pass) {
        set_last_error("log-mel plan needs " + std::to_string(pl->smem_bytes) + " B shared memory per CTA (limit " +
                       std::to_string(max_smem) + "): hop/n_fft combination too large");
        logmel_plan_destroy(pl);
        return 3;
    }
    *out = pl;
    return 0;
}

void logmel_plan_destroy(LogmelPlan* pl) {
    if (!pl) return;
    cudaFree(pl->hann);
    cudaFree(pl->hann_half);
    cudaFree(pl->twp);
    cudaFree(pl->mel_desc);
    cudaFree(pl->mel_w4);
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
    p.sub_frames = pl->sub_frames;
    p.tile_stride = pl->tile_stride;
    p.mel_nnz = pl->nnz;

    p.hann_half = pl->hann_half;
    p.twp = pl->twp;
    p.mel_desc = pl->mel_desc;
    p.mel_w4 = pl->mel_w4;
    p.mel_nnz4 = pl->nnz4;
    p.chunk_floats = pl->chunk_floats;
    p.chunk_bufs = pl->chunk_bufs;
    int cluster = pl->cluster;
    size_t smem_bytes = pl->smem_bytes;
    LogmelKernelFn fn = pick_logmel_kernel(pl->log2m, pl->smem_tables);
    if (pl->two_pass) {
        // cluster size for this launch: fewest (waves of resident clusters) x (frames a CTA walks)
        // (measured on the 240-window 48 kHz and 360-window 16 kHz workloads for clusters of 4 / 5 / 6 / 8: the ordering follows
        // waves x (frames per CTA + ~10 frames of per-CTA fixed cost), profiles/r2_logmel_two_pass.txt)
        const LogmelPlan::TwoPassCfg* best = nullptr;
        long long best_cost = 0;
        for (const auto& c : pl->tp) {
            const long long waves = ceil_div(n_win, c.resident_clusters);
            const long long cost = waves * (c.frames_per_cta + 10);
            if (best == nullptr || cost < best_cost) {
                best = &c;
                best_cost = cost;
            }
        }
        cluster = best->cluster;
        smem_bytes = best->smem_bytes;
        p.frames_per_cta = best->frames_per_cta;
        p.tile_stride = best->tile_stride;
        fn = pick_two_pass_kernel(pl->log2m, pl->chunk_bufs);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(n_win) * cluster);
    cfg.blockDim = dim3(kLogmelThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WSB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fn, p));
    count_launch();
    return 0;
}

size_t logmel_plan_smem(const LogmelPlan* pl) { return pl->smem_bytes; }
int logmel_plan_cluster(const LogmelPlan* pl) { return pl->cluster; }

}  // namespace wsb
