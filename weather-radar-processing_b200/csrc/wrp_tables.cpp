// wrp_tables.cpp — init-time constants of the chain, built on the host in double and
// rounded once.  Replaces generate_hamming_coefficients / generate_ma_coefficients
// (rpv2.cu:222-281) without FFTW: the N-point transform of the 7 zero-padded taps is a
// 7-term direct sum per bin.
#include "wrp_internal.h"

#include <cmath>

namespace wrp {

static const double kPi = 3.14159265358979323846;

void build_host_tables(int M, int N, int ma_taps, HostTables &t)
{
    // Window: wr(i) = 0.53836 - 0.46164 cos(2 pi i/(M-1)), wd(j) likewise; normalised by
    // the window power and by K = -1/(16383.5*M*N*sqrt(50))  (rpv2.cu:224-241).  The
    // accumulators are double as in the CPU oracle source read.cc:11-27.
    std::vector<double> wr(M), wd(N);
    double p_range = 0, p_doppler = 0;
    for (int i = 0; i < M; i++) {
        wr[i] = 0.53836 - 0.46164 * std::cos(2 * kPi * i / (M - 1));
        p_range += wr[i] * wr[i];
    }
    p_range /= M;
    for (int j = 0; j < N; j++) {
        wd[j] = 0.53836 - 0.46164 * std::cos(2 * kPi * j / (N - 1));
        p_doppler += wd[j] * wd[j];
    }
    p_doppler /= N;
    const double k_wind = -1.0 / (16383.5 * M * N * std::sqrt(50.0));
    t.c = k_wind / std::sqrt(p_range * p_doppler);

    t.ham.resize((size_t)M * N);
    t.wr_c.resize(M);
    t.wd.resize(N);
    for (int i = 0; i < M; i++) {
        t.wr_c[i] = (float)(wr[i] * t.c);
        for (int j = 0; j < N; j++) t.ham[(size_t)i * N + j] = (float)(wr[i] * wd[j] * t.c);
    }
    for (int j = 0; j < N; j++) t.wd[j] = (float)wd[j];

    // Moving-average taps: normalised Gaussian, centre (taps-1)/2 in integer arithmetic,
    // float accumulation as in rpv2.cu:254-262.
    t.taps.resize(ma_taps);
    float sum = 0.f;
    for (int i = 0; i < ma_taps; i++) {
        const double d = (double)(i - ((ma_taps - 1) / 2));
        t.taps[i] = (float)std::exp(-(d * d) / 2);
        sum += t.taps[i];
    }
    t.taps_sum = 0.f;
    for (int i = 0; i < ma_taps; i++) {
        t.taps[i] = t.taps[i] / sum;
        t.taps_sum += t.taps[i];
    }

    // fft_ma[k] = sum_t taps[t] exp(-2 pi i t k / N)   (rpv2.cu:264-275 via FFTW there)
    t.fft_ma.resize(2 * (size_t)N);
    for (int k = 0; k < N; k++) {
        double re = 0, im = 0;
        for (int q = 0; q < ma_taps && q < N; q++) {
            const double a = -2 * kPi * (double)((long long)q * k % N) / N;
            re += t.taps[q] * std::cos(a);
            im += t.taps[q] * std::sin(a);
        }
        t.fft_ma[2 * k] = (float)re;
        t.fft_ma[2 * k + 1] = (float)im;
    }

    t.tw_m.resize(2 * (size_t)M);
    for (int q = 0; q < M; q++) {
        const double a = -2 * kPi * q / M;
        t.tw_m[2 * q] = (float)std::cos(a);
        t.tw_m[2 * q + 1] = (float)std::sin(a);
    }
    t.tw_n.resize(2 * (size_t)N);
    for (int q = 0; q < N; q++) {
        const double a = 2 * kPi * q / N;
        t.tw_n[2 * q] = (float)std::cos(a);
        t.tw_n[2 * q + 1] = (float)std::sin(a);
    }
}

} // namespace wrp
