// CPU harness for whisperseg_b200/csrc/logmel_fft.cuh: runs the lane-level passes of the two-pass register FFT for every
// lane of one warp in program order (all lanes of a phase, then the next phase -- the warp barriers of the kernel) and
// compares the power spectrum with a float64 DFT.  Built and run by tests/test_logmel_fft_host.py (no GPU involved).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "logmel_fft.cuh"

using namespace wsb::lfft;

template <int LOG2M>
static double run(unsigned seed, int x_offset) {
    using T = TwoPass<LOG2M>;
    constexpr int M = T::M, N = 2 * M, G = T::G, PS = 16 * G + 1;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<float> hann_half(N);
    for (int i = 0; i < N; ++i) hann_half[i] = 0.5f * static_cast<float>(0.5 - 0.5 * std::cos(two_pi * i / N));
    std::vector<float2> tw(T::TWN), twp(T::TWP);
    for (int k = 0; k < T::TWN; ++k) tw[k] = make_float2((float)std::cos(two_pi * k / N), (float)-std::sin(two_pi * k / N));
    for (int k2 = 0; k2 < 16; ++k2)
        for (int n1 = 0; n1 < T::N1; ++n1)
            twp[k2 * T::N1 + n1] = make_float2((float)std::cos(two_pi * n1 * k2 / M), (float)-std::sin(two_pi * n1 * k2 / M));
    // G frames, hop 7 apart, starting at an odd or even float offset (exercises the unaligned load path)
    const int hop = 7 + (x_offset & 1);
    std::vector<float> samples(x_offset + N + hop * G + 8);
    srand(seed);
    for (auto& s : samples) s = (float)((rand() / (double)RAND_MAX) * 2.0 - 1.0);
    std::vector<float2> buf(T::BUF);
    std::vector<float> pt((M + 1) * PS, -1.0f);
    const int slot0 = 3;
    for (int lane = 0; lane < 32; ++lane) {
        const int f = T::frame_of(lane);
        const float* x = samples.data() + x_offset + f * hop;
        float2 h[16];
        T::load_window(lane, hann_half.data(), h);
        T::pass1(lane, x, ((x_offset + f * hop) & 1) == 0, h, twp.data(), buf.data());
    }
    float2 u[32][16];
    for (int lane = 0; lane < 32; ++lane) T::pass2_load(lane, buf.data(), u[lane]);
    for (int lane = 0; lane < 32; ++lane) T::pass2_store(lane, u[lane], buf.data());
    for (int lane = 0; lane < 32; ++lane) T::untangle(lane, buf.data(), tw.data(), pt.data(), PS, slot0, true);
    double worst = 0.0;
    for (int f = 0; f < G; ++f) {
        const float* x = samples.data() + x_offset + f * hop;
        double pmax = 0.0;
        std::vector<double> ref(M + 1);
        for (int k = 0; k <= M; ++k) {
            double re = 0.0, im = 0.0;
            for (int n = 0; n < N; ++n) {
                const double v = (double)x[n] * (0.5 - 0.5 * std::cos(two_pi * n / N));
                re += v * std::cos(two_pi * (double)((long long)k * n % N) / N);
                im -= v * std::sin(two_pi * (double)((long long)k * n % N) / N);
            }
            ref[k] = re * re + im * im;
            pmax = std::fmax(pmax, ref[k]);
        }
        for (int k = 0; k <= M; ++k) {
            const double got = pt[k * PS + slot0 + f];
            worst = std::fmax(worst, std::fabs(got - ref[k]) / pmax);
        }
    }
    // slots other than slot0 .. slot0+G-1 must be untouched
    for (int k = 0; k <= M; ++k)
        for (int s = 0; s < PS; ++s)
            if ((s < slot0 || s >= slot0 + G) && pt[k * PS + s] != -1.0f) return 1e9;
    return worst;
}

int main() {
    double w = 0.0;
    for (unsigned seed = 1; seed <= 3; ++seed)
        for (int off = 0; off < 2; ++off) {
            const double a = run<8>(seed, off), b = run<9>(seed, off);
            printf("seed %u offset %d: M=256 err %.3g  M=512 err %.3g\n", seed, off, a, b);
            w = std::fmax(w, std::fmax(a, b));
        }
    printf("worst %.3g\n", w);
    return w < 2e-6 ? 0 : 1;
}
