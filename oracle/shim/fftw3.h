/*
 * oracle/shim/fftw3.h — observing stand-in for FFTW so that the UNMODIFIED reference
 * sources read.cc / read_single.cc compile and run in this image (FFTW is not
 * installed and there is no network).  TEST INFRASTRUCTURE ONLY.
 *
 * It provides exactly the FFTW entry points the reference calls
 * (read.cc:87-98,153-154,188-189,278-280; read_single.cc:101-112,246-247,292-293,411-413)
 * with FFTW's conventions: FORWARD = exp(-2*pi*i*jk/n), BACKWARD un-normalised.
 * The transform itself is a radix-2 FFT on split arrays with per-stage twiddle tables
 * rounded from double, written here (deliberately a different code path from
 * oracle/wrp_oracle.c so the two check each other).
 *
 * Observation: when the environment variable WRP_SPY_DIR is set, every executed
 * transform appends (input[n], output[n]) to  $WRP_SPY_DIR/exec_<prec>_n<n>_<fwd|bwd>.bin
 * and every log10() call made by the translation unit appends its argument to
 * $WRP_SPY_DIR/log10_args.bin (as double).  The reference never prints its
 * products (its dump blocks are commented out), so this is how the tests see the
 * stage data of a literal reference run.
 */
#ifndef WRP_SHIM_FFTW3_H
#define WRP_SHIM_FFTW3_H

#include <math.h>
#include <cmath>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <map>
#include <vector>

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

typedef double fftw_complex[2];
typedef float fftwf_complex[2];

namespace wrp_shim {

inline const char *spy_dir() { return getenv("WRP_SPY_DIR"); }

inline FILE *spy_file(const std::string &name)
{
    static std::map<std::string, FILE *> files;
    auto it = files.find(name);
    if (it != files.end()) return it->second;
    std::string path = std::string(spy_dir()) + "/" + name;
    FILE *f = fopen(path.c_str(), "wb");
    files[name] = f;
    return f;
}

inline void spy_flush_all()
{
    fflush(NULL);
}

template <typename T> struct plan_t {
    int n;
    int sign;
    T (*in)[2];
    T (*out)[2];
    std::vector<T> twr, twi; /* per-stage twiddles, contiguous: stage `len` at offset len/2, from a double table */
    std::vector<int> rev;    /* bit-reversal permutation */
    std::vector<T> re, im;   /* work arrays (split real / imaginary) */
};

template <typename T> plan_t<T> *make_plan(int n, T (*in)[2], T (*out)[2], int sign)
{
    plan_t<T> *p = new plan_t<T>;
    p->n = n;
    p->sign = sign;
    p->in = in;
    p->out = out;
    /* stage len = 2, 4, ..., n uses W_len^k, k < len/2, stored at [len/2 + k]: the inner loops read
     * them contiguously (so the compiler vectorises them); computed in double, rounded once to T */
    p->twr.assign(n > 1 ? n : 2, (T)0);
    p->twi.assign(n > 1 ? n : 2, (T)0);
    for (int len = 2; len <= n; len <<= 1)
        for (int k = 0; k < len / 2; k++) {
            p->twr[len / 2 + k] = (T)cos(2.0 * M_PI * k / len);
            p->twi[len / 2 + k] = (T)(sign * sin(2.0 * M_PI * k / len));
        }
    int bits = 0;
    while ((1 << bits) < n) bits++;
    p->rev.resize(n);
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < bits; b++)
            if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        p->rev[i] = r;
    }
    p->re.resize(n);
    p->im.resize(n);
    return p;
}

/* iterative decimation-in-time radix-2 on split arrays, bit-reversed input order, computed in T.
 * The bit-reversal gather and the first two stages (twiddles 1 and -+i) are one pass over blocks of
 * four; the rest run contiguous, vectorisable inner loops.  About 4x the speed of the textbook loop it replaced
 * (this header is also what `bench.py --impl reference` times, so it should not be gratuitously
 * slow; real FFTW with SIMD codelets would still be ~2x faster on the transforms). */
template <typename T> void run_plan(plan_t<T> *p)
{
    const int n = p->n;
    T *__restrict__ re = p->re.data();
    T *__restrict__ im = p->im.data();
    static const bool spying = spy_dir() != NULL;
    FILE *f = NULL;
    if (spying) {
        char name[96];
        snprintf(name, sizeof name, "exec_%s_n%d_%s.bin", sizeof(T) == 8 ? "f64" : "f32", n,
                 p->sign < 0 ? "fwd" : "bwd");
        f = spy_file(name);
        fwrite(p->in, sizeof(T) * 2, n, f);
    }
    int len = 2;
    if (n >= 4) {
        /* bit-reversal gather fused with stages len = 2 and len = 4: the four inputs of output block
         * `base` sit at rev[base] + {0, n/2, n/4, 3n/4}; W_4^1 = -i (forward) or +i (backward) */
        const T s = (T)p->sign;
        const int *__restrict__ rev = p->rev.data();
        const T(*__restrict__ x)[2] = p->in;
        const int q = n / 4;
        for (int base = 0; base < n; base += 4) {
            const int r = rev[base];
            const T x0r = x[r][0], x0i = x[r][1], x1r = x[r + 2 * q][0], x1i = x[r + 2 * q][1];
            const T x2r = x[r + q][0], x2i = x[r + q][1], x3r = x[r + 3 * q][0], x3i = x[r + 3 * q][1];
            const T a0r = x0r + x1r, a0i = x0i + x1i, a1r = x0r - x1r, a1i = x0i - x1i;
            const T b0r = x2r + x3r, b0i = x2i + x3i, b1r = x2r - x3r, b1i = x2i - x3i;
            const T tr = -s * b1i, ti = s * b1r; /* b1 * (s i) */
            re[base] = a0r + b0r;
            im[base] = a0i + b0i;
            re[base + 1] = a1r + tr;
            im[base + 1] = a1i + ti;
            re[base + 2] = a0r - b0r;
            im[base + 2] = a0i - b0i;
            re[base + 3] = a1r - tr;
            im[base + 3] = a1i - ti;
        }
        len = 8;
    } else {
        for (int i = 0; i < n; i++) {
            re[p->rev[i]] = p->in[i][0];
            im[p->rev[i]] = p->in[i][1];
        }
    }
    for (; len <= n; len <<= 1) {
        const int half = len / 2;
        const T *__restrict__ wr = p->twr.data() + half;
        const T *__restrict__ wi = p->twi.data() + half;
        for (int base = 0; base < n; base += len) {
            T *__restrict__ ar = re + base, *__restrict__ ai = im + base;
            T *__restrict__ br = re + base + half, *__restrict__ bi = im + base + half;
#pragma GCC ivdep
            for (int k = 0; k < half; k++) {
                const T tr = br[k] * wr[k] - bi[k] * wi[k];
                const T ti = br[k] * wi[k] + bi[k] * wr[k];
                br[k] = ar[k] - tr;
                bi[k] = ai[k] - ti;
                ar[k] = ar[k] + tr;
                ai[k] = ai[k] + ti;
            }
        }
    }
    for (int i = 0; i < n; i++) {
        p->out[i][0] = re[i];
        p->out[i][1] = im[i];
    }
    if (f) fwrite(p->out, sizeof(T) * 2, n, f);
}

inline double spy_log10(double v)
{
    if (spy_dir()) {
        FILE *f = spy_file("log10_args.bin");
        fwrite(&v, sizeof v, 1, f);
    }
    return ::log10(v);
}
inline float spy_log10(float v)
{
    if (spy_dir()) {
        double d = v;
        FILE *f = spy_file("log10_args.bin");
        fwrite(&d, sizeof d, 1, f);
    }
    return ::log10f(v);
}

} // namespace wrp_shim

typedef wrp_shim::plan_t<double> *fftw_plan;
typedef wrp_shim::plan_t<float> *fftwf_plan;

inline void *fftw_malloc(size_t n) { return malloc(n); }
inline void fftw_free(void *p) { free(p); }
inline fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned)
{
    return wrp_shim::make_plan<double>(n, in, out, sign);
}
inline void fftw_execute(fftw_plan p) { wrp_shim::run_plan(p); }
inline void fftw_destroy_plan(fftw_plan p) { delete p; }

inline void *fftwf_malloc(size_t n) { return malloc(n); }
inline void fftwf_free(void *p) { free(p); }
inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned)
{
    return wrp_shim::make_plan<float>(n, in, out, sign);
}
inline void fftwf_execute(fftwf_plan p) { wrp_shim::run_plan(p); }
inline void fftwf_destroy_plan(fftwf_plan p) { delete p; }

/* observe the dB stage (read.cc:342-343, read_single.cc:469-470) */
/* (not under nvcc: the reference's GPU variants call log10 in device code, and only need this
 * header for the init-time transform of the 7 moving-average taps, e.g. gpu_1fp_unistream.cu:405-416) */
#ifndef __CUDACC__
#define log10(x) wrp_shim::spy_log10(x)
#endif

#endif
