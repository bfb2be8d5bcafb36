// K1 building blocks: the two-pass register FFT of the log-mel kernel, lane by lane.
//
// A frame's real FFT of n_fft = 2M points is an M-point complex FFT of z[j] = x[2j] + i x[2j+1] plus an untangle
// (HF feature_extraction_whisper.py:135-164 calls torch.stft; this is the same transform).  M = N1 x 16:
//
//   pass 1   lane n1 holds z[n1 + N1 n2], n2 = 0..15: a 16-point DIF in registers gives Y[n1][k2], which is multiplied by
//            W_M^(n1 k2) and written to the warp's shared-memory buffer at [n1][k2] (row stride 17: conflict-free both ways)
//   pass 2   Z[16 k1 + k2] = sum_n1 Y'[n1][k2] W_N1^(n1 k1):
//              M = 256 (N1 = 16): lane (frame, k2) reads its column and runs the same 16-point DIF; a warp carries 2 frames
//              M = 512 (N1 = 32): lane (k2, half) forms the first radix-2 stage of the 32-point DIF straight from the buffer
//                                 (half 0: a + b -> even k1, half 1: (a - b) W_32^n1 -> odd k1) and runs the 16-point DIF
//            the spectrum goes back to the same buffer in natural order, 32 consecutive float2 per store instruction
//   untangle one thread per pair (k, M - k): X[k] = E + W_2M^k O, X[M-k] = conj(E - W_2M^k O); both powers are written to the
//            CTA-wide transposed array P[k][frame] that the mel stage reads with lane = frame
//
// The Hann window handed in is pre-multiplied by 0.5 (exact), which absorbs the two 1/2 factors of E and O.
//
// Everything here is __host__ __device__ so that the index math runs on the CPU too
// (tests/test_logmel_fft_host.py compiles tests/host/logmel_fft_host.cpp and compares with a float64 DFT).
#pragma once
#include <cuda_runtime.h>

namespace wsb {
namespace lfft {

#define WSB_HD __host__ __device__ __forceinline__

WSB_HD float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

WSB_HD constexpr int bitrev4(int v) { return ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3); }

// d * exp(-2 pi i k / 16), k = 0..7 (k is a compile-time constant after unrolling)
WSB_HD float2 rot16(float2 d, int k) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
    switch (k) {
        case 0: return d;
        case 1: return make_float2(d.x * C1 + d.y * S1, d.y * C1 - d.x * S1);
        case 2: return make_float2((d.x + d.y) * R, (d.y - d.x) * R);
        case 3: return make_float2(d.x * S1 + d.y * C1, d.y * S1 - d.x * C1);
        case 4: return make_float2(d.y, -d.x);
        case 5: return make_float2(d.y * C1 - d.x * S1, -d.y * S1 - d.x * C1);
        case 6: return make_float2((d.y - d.x) * R, -(d.x + d.y) * R);
        default: return make_float2(d.y * S1 - d.x * C1, -d.y * C1 - d.x * S1);
    }
}

// 16-point radix-2 DIF in registers: v[i] <- X[bitrev4(i)]
WSB_HD void fft16(float2 (&v)[16]) {
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((i & h) == 0) {
                const float2 a = v[i], b = v[i + h];
                v[i] = make_float2(a.x + b.x, a.y + b.y);
                v[i + h] = rot16(make_float2(a.x - b.x, a.y - b.y), (i & (h - 1)) * (8 / h));
            }
        }
    }
}

// d * exp(-2 pi i n / 32), n = 0..15 (compile-time n)
WSB_HD float2 rot32(float2 d, int n) {
    constexpr float C[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                             0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
                             -0.38268343236508977f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                             -0.92387953251128674f, -0.98078528040323043f};
    constexpr float S[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f, 0.70710678118654752f,
                             0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.0f, 0.98078528040323043f,
                             0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                             0.38268343236508977f, 0.19509032201612825f};
    if (n == 0) return d;
    if (n == 8) return make_float2(d.y, -d.x);
    return make_float2(d.x * C[n] + d.y * S[n], d.y * C[n] - d.x * S[n]);
}

template <int LOG2M>
struct TwoPass {
    static_assert(LOG2M == 8 || LOG2M == 9, "two-pass register FFT: M = 256 or 512");
    static constexpr int M = 1 << LOG2M;
    static constexpr int N1 = M / 16;             // lanes that share a frame in pass 1 (16 or 32)
    static constexpr int G = 32 / N1;             // frames a warp carries per iteration (2 or 1)
    static constexpr int ROW = 17;                // float2 per n1 row of the pass-1 -> pass-2 buffer
    static constexpr int FRAME = 16 * ROW;        // float2 between the two frames of a warp (M = 256)
    static constexpr int BUF = 32 * ROW;          // float2 per warp (544): M = 512 uses 32 rows, M = 256 uses 2 x 16
    static constexpr int TWP = 16 * N1;           // pass-1 twiddle table entries, [k2][n1]
    static constexpr int TWN = M / 2 + 1;         // untangle twiddles W_2M^k, k = 0..M/2

    static WSB_HD int frame_of(int lane) { return G == 2 ? (lane >> 4) : 0; }

    // pass 1: x = the frame's first sample (shared memory), hann_half = 0.5 * periodic Hann, twp = [k2][n1] twiddles
    // the 16 window pairs a lane needs are the same for every frame it touches: hann_half[2j], hann_half[2j+1], j = n1 + N1 n2
    static WSB_HD void load_window(int lane, const float* hann_half, float2 (&h)[16]) {
        const int n1 = lane & (N1 - 1);
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) h[n2] = reinterpret_cast<const float2*>(hann_half)[n1 + N1 * n2];
    }

    static WSB_HD void pass1(int lane, const float* x, bool x_aligned, const float2 (&h)[16], const float2* twp, float2* buf) {
        const int n1 = lane & (N1 - 1);
        float2 v[16];
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            const int j = n1 + N1 * n2;
            const float2 xv = x_aligned ? reinterpret_cast<const float2*>(x)[j] : make_float2(x[2 * j], x[2 * j + 1]);
            v[n2] = make_float2(xv.x * h[n2].x, xv.y * h[n2].y);
        }
        fft16(v);
        float2* row = buf + frame_of(lane) * FRAME + n1 * ROW;
        row[0] = v[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            const int k2 = bitrev4(i);
            row[k2] = cmulf(v[i], twp[k2 * N1 + n1]);
        }
    }

    // pass 2, first half: read the column(s) and transform; u[i] = Z[index2(lane, i)]
    static WSB_HD void pass2_load(int lane, const float2* buf, float2 (&u)[16]) {
        const int k2 = lane & 15;
        if (G == 2) {
            const float2* col = buf + (lane >> 4) * FRAME + k2;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) u[n1] = col[n1 * ROW];
        } else {
            const float2* col = buf + k2;
            const bool odd = (lane >> 4) != 0;
            const float sgn = odd ? -1.0f : 1.0f;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const float2 a = col[n1 * ROW], b = col[(n1 + 16) * ROW];
                u[n1] = make_float2(a.x + sgn * b.x, a.y + sgn * b.y);
            }
            if (odd) {
#pragma unroll
                for (int n1 = 1; n1 < 16; ++n1) u[n1] = rot32(u[n1], n1);
            }
        }
        fft16(u);
    }

    // pass 2, second half (after a warp barrier: the spectrum overwrites the buffer): natural order, frame f at f * M
    static WSB_HD void pass2_store(int lane, const float2 (&u)[16], float2* buf) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int q = bitrev4(i);
            if (G == 2) buf[(lane >> 4) * M + 16 * q + (lane & 15)] = u[i];
            else buf[32 * q + lane] = u[i];
        }
    }

    // untangle + power: P[k][slot] for k = 0..M; tw[k] = exp(-2 pi i k / 2M); pt row stride ps; slot0 = the warp's
    // first frame slot of this iteration (frame f of the warp goes to slot0 + f)
    static WSB_HD void untangle(int lane, const float2* buf, const float2* tw, float* pt, int ps, int slot0, bool active) {
        if (!active) return;
        const int f = frame_of(lane);
        const float2* z = buf + f * M;
        float* col = pt + slot0 + f;
        const int kk = lane & (N1 - 1);
#pragma unroll
        for (int i = 0; i < M / 2 / N1; ++i) {             // pairs (k, M - k), k = 0 .. M/2 - 1 (k = 0 pairs with itself: P[0], P[M])
            const int k = kk + i * N1;
            const float2 zk = z[k], zm = z[(M - k) & (M - 1)];
            const float2 e = make_float2(zk.x + zm.x, zk.y - zm.y);
            const float2 o = make_float2(zk.y + zm.y, zm.x - zk.x);
            const float2 t = cmulf(o, tw[k]);
            const float ax = e.x + t.x, ay = e.y + t.y, bx = e.x - t.x, by = e.y - t.y;
            col[k * ps] = ax * ax + ay * ay;
            col[(M - k) * ps] = bx * bx + by * by;
        }
        if (kk == N1 - 1) {                                 // the self-paired middle bin: X[M/2] = 2 conj(Z[M/2])
            const float2 zc = z[M / 2];
            col[(M / 2) * ps] = 4.0f * (zc.x * zc.x + zc.y * zc.y);
        }
    }
};

}  // namespace lfft
}  // namespace wsb
