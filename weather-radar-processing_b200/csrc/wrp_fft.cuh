// wrp_fft.cuh — in-register radix-R FFT building blocks (R = 2..32) for sm_100a.
//
// Everything is a compile-time-unrolled decimation-in-time network on a float2
// register array: twiddles are immediates, every array index is a constant, so the
// array lives in registers and ptxas sees straight-line FFMA/FADD code.  The
// butterfly uses the FMA form  a' = a + w*b (4 FFMA), b' = 2a - a' (2 FFMA): six
// FMA-pipe instructions per butterfly instead of eight.  Outputs that the caller
// never reads are dead-code-eliminated, which is how the pruned range FFT (only
// rows k < M/2 survive stage 04, rpv2.cu:502) drops half of its last stage.
//
// Replaces the cuFFT plans of the reference (rpv2.cu:318-341); cuFFT is not used.
#pragma once

#include <cuda_runtime.h>
#include <type_traits>
#include <utility>

namespace wrp {

// cos/sin(2*pi*t/32), t = 0..16, rounded once from double.
__host__ __device__ constexpr float cos32(int t)
{
    constexpr float tab[17] = {1.f,
                               0.98078528040323043f,
                               0.92387953251128674f,
                               0.83146961230254524f,
                               0.70710678118654757f,
                               0.55557023301960229f,
                               0.38268343236508984f,
                               0.19509032201612833f,
                               0.f,
                               -0.19509032201612819f,
                               -0.38268343236508973f,
                               -0.55557023301960196f,
                               -0.70710678118654746f,
                               -0.83146961230254535f,
                               -0.92387953251128674f,
                               -0.98078528040323043f,
                               -1.f};
    return tab[t];
}
__host__ __device__ constexpr float sin32(int t)
{
    constexpr float tab[17] = {0.f,
                               0.19509032201612825f,
                               0.38268343236508978f,
                               0.55557023301960218f,
                               0.70710678118654746f,
                               0.83146961230254524f,
                               0.92387953251128674f,
                               0.98078528040323043f,
                               1.f,
                               0.98078528040323043f,
                               0.92387953251128674f,
                               0.83146961230254546f,
                               0.70710678118654757f,
                               0.55557023301960218f,
                               0.38268343236508989f,
                               0.19509032201612861f,
                               0.f};
    return tab[t];
}

// compile-time loop: f(std::integral_constant<int, I>{}) for I in [0, N)
template <typename F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F> __device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

template <int R> __host__ __device__ constexpr int log2c()
{
    int l = 0;
    for (int r = R; r > 1; r >>= 1) ++l;
    return l;
}
// bit reversal of i within log2(R) bits
template <int R> __host__ __device__ constexpr int brev(int i)
{
    int r = 0;
    for (int b = 0; b < log2c<R>(); ++b)
        if (i & (1 << b)) r |= 1 << (log2c<R>() - 1 - b);
    return r;
}

// Packed complex add/sub: one sm_100 FADD2 on the (re, im) register pair instead of two FADDs.
// Same FP32 throughput (measured: tools/micro/ffma2.cu), half the issue slots — and the chain is
// issue-bound.  ptxas keeps float2 values in aligned pairs, so no moves are added.
__device__ __forceinline__ float2 cadd(float2 a, float2 b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b)
{
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&d);
}

__device__ __forceinline__ float2 cmul2(float2 a, float2 b) // element-wise (a.x*b.x, a.y*b.y), one FMUL2
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 cfma2(float2 a, float2 b, float2 c) // element-wise a*b + c, one FFMA2
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
          "l"(*reinterpret_cast<unsigned long long *>(&c)));
    return *reinterpret_cast<float2 *>(&d);
}

// One DIT butterfly with twiddle w = exp(SIGN * 2*pi*i * T/32): (a, b) <- (a + w b, a - w b).
template <int T, int SIGN> __device__ __forceinline__ void bfly(float2 &a, float2 &b)
{
    if constexpr (T == 0) {
        const float2 t = b;
        b = csub(a, t);
        a = cadd(a, t);
    } else if constexpr (T == 8) {
        // w = SIGN * i
        const float2 t = SIGN > 0 ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
        b = make_float2(a.x - t.x, a.y - t.y);
        a = make_float2(a.x + t.x, a.y + t.y);
    } else {
        constexpr float wr = cos32(T);
        constexpr float wi = SIGN * sin32(T);
        float tx = fmaf(wr, b.x, a.x);
        tx = fmaf(-wi, b.y, tx);
        float ty = fmaf(wr, b.y, a.y);
        ty = fmaf(wi, b.x, ty);
        b.x = fmaf(2.f, a.x, -tx);
        b.y = fmaf(2.f, a.y, -ty);
        a.x = tx;
        a.y = ty;
    }
}

// One DIT stage of span H on R points.
template <int R, int H, int SIGN> __device__ __forceinline__ void dit_stage(float2 (&v)[R])
{
    static_for<R / 2>([&](auto idx) {
        constexpr int q = decltype(idx)::value;
        constexpr int k = q % H;            // position inside the half block
        constexpr int base = (q / H) * 2 * H; // block start
        constexpr int T = k * (16 / H);     // W_{2H}^k in units of 2*pi/32
        bfly<T, SIGN>(v[base + k], v[base + k + H]);
    });
}

// In-place radix-R DIT FFT.  INPUT must sit in bit-reversed slots: v[brev<R>(n)] = x[n].
// OUTPUT is in natural order: v[k] = sum_n x[n] exp(SIGN*2*pi*i*n*k/R).
template <int R, int SIGN> __device__ __forceinline__ void fft_dit(float2 (&v)[R])
{
    static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
    dit_stage<R, 1, SIGN>(v);
    if constexpr (R >= 4) dit_stage<R, 2, SIGN>(v);
    if constexpr (R >= 8) dit_stage<R, 4, SIGN>(v);
    if constexpr (R >= 16) dit_stage<R, 8, SIGN>(v);
    if constexpr (R >= 32) dit_stage<R, 16, SIGN>(v);
}

// Same network without its first stage (span 1): the caller has already formed
// (v[2m], v[2m+1]) <- (v[2m] + v[2m+1], v[2m] - v[2m+1]), e.g. fused with a window multiply.
template <int R, int SIGN> __device__ __forceinline__ void fft_dit_after_stage1(float2 (&v)[R])
{
    static_assert(R == 4 || R == 8 || R == 16 || R == 32, "radix");
    dit_stage<R, 2, SIGN>(v);
    if constexpr (R >= 8) dit_stage<R, 4, SIGN>(v);
    if constexpr (R >= 16) dit_stage<R, 8, SIGN>(v);
    if constexpr (R >= 32) dit_stage<R, 16, SIGN>(v);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}

// acc += g * exp(-2*pi*i * T32/32), constant twiddle (4 FFMA)
template <int T32> __device__ __forceinline__ void cmac_w32(float2 &acc, float2 g)
{
    constexpr float wr = cos32(T32), wi = -sin32(T32);
    acc.x = fmaf(g.x, wr, acc.x);
    acc.x = fmaf(-g.y, wi, acc.x);
    acc.y = fmaf(g.x, wi, acc.y);
    acc.y = fmaf(g.y, wr, acc.y);
}

// Bins 0, 1 and 2 of the forward R-point DFT  X[q] = sum_a x[a] exp(-2*pi*i*a*q/R)  of a register
// array in NATURAL order (R = 16 or 32) — a decimation-in-frequency network pruned to the three
// outputs the energy form of the Doppler stage needs (wrp_persistent.cu):
//   s = x[a] + x[a+R/2], d = x[a] - x[a+R/2];   X[1] = sum_{a<R/4} (d[a] - i d[a+R/4]) W_R^a
//   ss = s[a] + s[a+R/4], sd = s[a] - s[a+R/4]; X[0] = sum ss;  X[2] = sum_{a<R/8} (sd[a] - i sd[a+R/8]) W_{R/2}^a
// About 3.5 instructions per input point, against ~23 for the full two-pass transform.
template <int R>
__device__ __forceinline__ void dft_bins012(const float2 (&x)[R], float2 &b0, float2 &b1, float2 &b2)
{
    static_assert(R == 16 || R == 32, "radix");
    float2 s[R / 2], d[R / 2];
    static_for<R / 2>([&](auto ai) {
        constexpr int a = decltype(ai)::value;
        s[a] = cadd(x[a], x[a + R / 2]);
        d[a] = csub(x[a], x[a + R / 2]);
    });
    float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    static_for<R / 4>([&](auto ai) {
        constexpr int a = decltype(ai)::value;
        const float2 g = make_float2(d[a].x + d[a + R / 4].y, d[a].y - d[a + R / 4].x);
        if constexpr (a == 0) acc[0] = g;
        else cmac_w32<a * (32 / R)>(acc[a & 1], g);
    });
    b1 = cadd(acc[0], acc[1]);
    float2 ss[R / 4], sd[R / 4];
    static_for<R / 4>([&](auto ai) {
        constexpr int a = decltype(ai)::value;
        ss[a] = cadd(s[a], s[a + R / 4]);
        sd[a] = csub(s[a], s[a + R / 4]);
    });
    static_for<R / 8>([&](auto ai) { // pairwise tree over ss
        constexpr int a = decltype(ai)::value;
        ss[a] = cadd(ss[a], ss[a + R / 8]);
    });
    if constexpr (R == 32) {
        ss[0] = cadd(ss[0], ss[2]);
        ss[1] = cadd(ss[1], ss[3]);
    }
    b0 = cadd(ss[0], ss[1]);
    float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    static_for<R / 8>([&](auto ai) {
        constexpr int a = decltype(ai)::value;
        const float2 g = make_float2(sd[a].x + sd[a + R / 8].y, sd[a].y - sd[a + R / 8].x);
        if constexpr (a == 0) acc2[0] = g;
        else cmac_w32<a * (64 / R)>(acc2[a & 1], g);
    });
    b2 = cadd(acc2[0], acc2[1]);
}

} // namespace wrp
