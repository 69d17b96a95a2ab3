/*
 * oracle/shim/fftw3.h — observing stand-in for FFTW so that the UNMODIFIED reference
 * sources read.cc / read_single.cc compile and run in this image (FFTW is not
 * installed and there is no network).  TEST INFRASTRUCTURE ONLY.
 *
 * It provides exactly the FFTW entry points the reference calls
 * (read.cc:87-98,153-154,188-189,278-280; read_single.cc:101-112,246-247,292-293,411-413)
 * with FFTW's conventions: FORWARD = exp(-2*pi*i*jk/n), BACKWARD un-normalised.
 * The transform itself is a plain double-precision-twiddle radix-2 FFT written here
 * (deliberately a different code path from oracle/wrp_oracle.c so the two check
 * each other).
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
    std::vector<double> cs; /* cos/sin table, double */
    std::vector<T> wr, wi;  /* the same table rounded once to T */
    std::vector<int> rev;   /* bit-reversal permutation */
    std::vector<T> re, im;  /* work arrays */
};

template <typename T> plan_t<T> *make_plan(int n, T (*in)[2], T (*out)[2], int sign)
{
    plan_t<T> *p = new plan_t<T>;
    p->n = n;
    p->sign = sign;
    p->in = in;
    p->out = out;
    p->cs.resize(2 * (size_t)n);
    for (int t = 0; t < n; t++) {
        p->cs[2 * t] = cos(2.0 * M_PI * t / n);
        p->cs[2 * t + 1] = sign * sin(2.0 * M_PI * t / n);
    }
    p->wr.resize(n);
    p->wi.resize(n);
    for (int t = 0; t < n; t++) {
        p->wr[t] = (T)p->cs[2 * t];
        p->wi[t] = (T)p->cs[2 * t + 1];
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

/* iterative decimation-in-time radix-2, bit-reversed input order, computed in T */
template <typename T> void run_plan(plan_t<T> *p)
{
    const int n = p->n;
    T *re = p->re.data(), *im = p->im.data();
    for (int i = 0; i < n; i++) {
        const int r = p->rev[i];
        re[r] = p->in[i][0];
        im[r] = p->in[i][1];
    }
    const bool spying = spy_dir() != NULL;
    FILE *f = NULL;
    if (spying) {
        char name[96];
        snprintf(name, sizeof name, "exec_%s_n%d_%s.bin", sizeof(T) == 8 ? "f64" : "f32", n,
                 p->sign < 0 ? "fwd" : "bwd");
        f = spy_file(name);
        fwrite(p->in, sizeof(T) * 2, n, f);
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len / 2, step = n / len;
        for (int base = 0; base < n; base += len) {
            for (int k = 0; k < half; k++) {
                const T wr = p->wr[k * step], wi = p->wi[k * step];
                const int a = base + k, b = a + half;
                const T tr = re[b] * wr - im[b] * wi;
                const T ti = re[b] * wi + im[b] * wr;
                re[b] = re[a] - tr;
                im[b] = im[a] - ti;
                re[a] = re[a] + tr;
                im[a] = im[a] + ti;
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
#define log10(x) wrp_shim::spy_log10(x)

#endif
