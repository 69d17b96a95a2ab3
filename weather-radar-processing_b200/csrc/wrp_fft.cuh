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

} // namespace wrp
