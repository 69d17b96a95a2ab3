// wrp_fused.cu — v1 of the fused path (two kernels per chunk of sectors) and the wire decoder.
//
// The product path is the persistent kernel in wrp_persistent.cu; this two-kernel form is kept for
// A/B measurements (WRP_FUSED_IMPL=v1: 109 k sectors/s against 216 k) and as the simplest statement
// of the two fused stages:
//
//   range_fft_kernel   stage 01 (window on load) + stage 02 (range FFT along i), writing
//                      only the rows k < M/2 that survive stage 04 (rpv2.cu:502).
//                      Replaces __apply_hamming + the strided cuFFT plan
//                      (rpv2.cu:86-91, 318-333, 418-428).
//   doppler_kernel     per-row mean removal + conj/FFT/conj (= un-normalised inverse DFT)
//                      + fftshift + clip (stage 03), |.|^2 (04), moving-average power
//                      (05-08 collapsed: the circular convolution's row sum is
//                      (sum of taps) * row sum), row power and the
//                      ZdB/ZDR products (09/10).  Replaces __sum_v4, __avgconj, cuFFT
//                      Doppler, __conjugate, __shift, __clip_v2, __abssqr, the pdop
//                      FFT pair, __apply_ma, __scale_real, __sum_inplace_v4 and
//                      __calcresult_v2 (rpv2.cu:93-213, 434-566).
//   decode_wire_kernel wire records -> planar complex float (sector.cpp:52-62 +
//                      rpv2.cu:369-383) for WRP_FMT_WIRE_I16BE input (used by both forms).
//
// Both FFT kernels are two-pass Cooley-Tukey with in-register radix-16/32 passes
// (wrp_fft.cuh) and one shared-memory exchange; loads are straight global->register,
// coalesced in 64-byte (range: 8 columns x c64) or 256-byte (Doppler: a warp per row
// segment) pieces.
#include "wrp_fft.cuh"
#include "wrp_internal.h"

namespace wrp {

__device__ __forceinline__ float2 ld_stream(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// --------------------------------------------------------------------------------------
// Kernel A: range FFT of T adjacent columns of one (sector, channel) plane.
//   M = R1*R2.  pass 1: thread (c, b) transforms the R1 samples i = R2*a + b (a < R1);
//   twiddle by exp(-2 pi i b ka / M); exchange through shared memory;
//   pass 2: thread (c, ka) transforms over b and writes rows k = ka + R1*kb.
// --------------------------------------------------------------------------------------
template <int R1, int R2, int T> struct RangeCfg {
    static constexpr int M = R1 * R2;
    static constexpr int THREADS = T * R2;
    static constexpr int ROW = R1 + 2;           // float2 units; +16 B keeps STS.128 conflict-free
    static constexpr int COL = R2 * ROW + 2;     // +16 B keeps the column planes on distinct banks
    static constexpr size_t SMEM = (size_t)T * COL * sizeof(float2);
};

template <int R1, int R2, int T, bool FULL>
__global__ void __launch_bounds__(T *R2, 2)
    range_fft_kernel(const float2 *__restrict__ iq, float2 *__restrict__ x2, const float *__restrict__ wrc_t,
                     const float *__restrict__ wd, const float2 *__restrict__ tw_a, int N)
{
    static_assert(R1 == R2, "pass-1 and pass-2 thread shapes are shared");
    using Cfg = RangeCfg<R1, R2, T>;
    constexpr int M = Cfg::M;
    extern __shared__ __align__(16) float2 ex[];

    const int c = threadIdx.x % T;
    const int b = threadIdx.x / T;
    const int col = blockIdx.x * T + c;
    const size_t plane = blockIdx.y;

    float2 v[R1];
    {
        const float2 *in = iq + plane * (size_t)M * N + col;
        static_for<R1>([&](auto ai) {
            constexpr int a = decltype(ai)::value;
            v[brev<R1>(a)] = ld_stream(in + (size_t)(R2 * a + b) * N);
        });
        // stage 01: x *= wr(i)*c * wd(j)   (rpv2.cu:86-91 with ham = wr*wd*c, :245-249)
        const float wdj = __ldg(wd + col);
        const float4 *w4 = reinterpret_cast<const float4 *>(wrc_t) + b * (R1 / 4);
        static_for<R1 / 4>([&](auto qi) {
            constexpr int q = decltype(qi)::value;
            const float4 w = __ldg(w4 + q);
            const float w0 = w.x * wdj, w1 = w.y * wdj, w2 = w.z * wdj, w3 = w.w * wdj;
            v[brev<R1>(4 * q + 0)].x *= w0;
            v[brev<R1>(4 * q + 0)].y *= w0;
            v[brev<R1>(4 * q + 1)].x *= w1;
            v[brev<R1>(4 * q + 1)].y *= w1;
            v[brev<R1>(4 * q + 2)].x *= w2;
            v[brev<R1>(4 * q + 2)].y *= w2;
            v[brev<R1>(4 * q + 3)].x *= w3;
            v[brev<R1>(4 * q + 3)].y *= w3;
        });
    }
    fft_dit<R1, -1>(v);
    {
        // inter-pass twiddle exp(-2 pi i b ka / M), two per 128-bit load
        const float4 *t4 = reinterpret_cast<const float4 *>(tw_a) + b * (R1 / 2);
        float2 *dst = ex + c * Cfg::COL + b * Cfg::ROW;
        static_for<R1 / 2>([&](auto qi) {
            constexpr int q = decltype(qi)::value;
            const float4 w = __ldg(t4 + q);
            const float2 y0 = q == 0 ? v[0] : cmul(v[2 * q], make_float2(w.x, w.y));
            const float2 y1 = cmul(v[2 * q + 1], make_float2(w.z, w.w));
            *reinterpret_cast<float4 *>(dst + 2 * q) = make_float4(y0.x, y0.y, y1.x, y1.y);
        });
    }
    __syncthreads();
    {
        const int ka = b;
        const float2 *src = ex + c * Cfg::COL + ka;
        static_for<R2>([&](auto bi) {
            constexpr int bb = decltype(bi)::value;
            v[brev<R2>(bb)] = src[bb * Cfg::ROW];
        });
        fft_dit<R2, -1>(v);
        constexpr int KEEP = FULL ? R2 : R2 / 2; // rows k < M/2 <=> kb < R2/2
        constexpr int ROWS_OUT = FULL ? M : M / 2;
        float2 *out = x2 + plane * (size_t)ROWS_OUT * N + col;
        static_for<KEEP>([&](auto ki) {
            constexpr int kb = decltype(ki)::value;
            out[(size_t)(ka + R1 * kb) * N] = v[kb];
        });
    }
}

// --------------------------------------------------------------------------------------
// Kernel B: Doppler transform + epilogue on rows of the range-FFT output.
//   N = R1*32.  A CTA owns ROWS = 256/R1 rows: either ROWS/2 gates x (hh, vv), or ROWS
//   gates of a single channel (vh, or hh when n_channels == 1).
//   pass 1: thread (row, lane l) transforms the R1 samples j = 32 a + l; the row mean is
//   removed on the a-sum (bin ka = 0) after a warp reduction; twiddle exp(+2 pi i l ka/N);
//   pass 2: thread (row, ka) transforms over l; bins ka + R1*kb.
// --------------------------------------------------------------------------------------
template <int R1> struct DopplerCfg {
    static constexpr int N = R1 * 32;
    static constexpr int THREADS = 256;
    static constexpr int ROWS = THREADS / R1;
    static constexpr int ROWS_PER_WARP = ROWS / 8;
    static constexpr int LROW = R1 + 2; // float2 units
    static constexpr int RROW = 32 * LROW;
    static constexpr size_t SMEM = (size_t)ROWS * RROW * sizeof(float2);
};

template <int R1>
__global__ void __launch_bounds__(256, 2)
    doppler_kernel(const float2 *__restrict__ x2, float *__restrict__ out, float *__restrict__ power,
                   const float2 *__restrict__ tw_b, int half_m, int C, int pair_blocks, float range_res,
                   float calib, float taps_sum)
{
    using Cfg = DopplerCfg<R1>;
    constexpr int N = Cfg::N;
    constexpr int ROWS = Cfg::ROWS;
    extern __shared__ __align__(16) float2 ex[];
    __shared__ float p_row[ROWS];

    const int sector = blockIdx.y;
    const bool pair = (int)blockIdx.x < pair_blocks;
    // row r of this CTA -> (channel, gate)
    auto row_channel = [&](int r) { return pair ? (r & 1) : (C == 1 ? 0 : 2); };
    auto row_gate = [&](int r) {
        return pair ? (int)blockIdx.x * (ROWS / 2) + (r >> 1) : ((int)blockIdx.x - pair_blocks) * ROWS + r;
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        // twiddles exp(+2 pi i lane ka / N), ka < R1, reused for every row of this thread
        float2 tw[R1];
        const float4 *t4 = reinterpret_cast<const float4 *>(tw_b) + lane * (R1 / 2);
        static_for<R1 / 2>([&](auto qi) {
            constexpr int q = decltype(qi)::value;
            const float4 w = __ldg(t4 + q);
            tw[2 * q] = make_float2(w.x, w.y);
            tw[2 * q + 1] = make_float2(w.z, w.w);
        });
#pragma unroll
        for (int rr = 0; rr < Cfg::ROWS_PER_WARP; ++rr) {
            const int r = warp * Cfg::ROWS_PER_WARP + rr;
            const float2 *in =
                x2 + (((size_t)sector * C + row_channel(r)) * half_m + row_gate(r)) * (size_t)N + lane;
            float2 v[R1];
            static_for<R1>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                v[brev<R1>(a)] = __ldcg(in + 32 * a);
            });
            fft_dit<R1, +1>(v);
            // mean removal (rpv2.cu:93-130): only the a-sum (ka = 0) carries the row mean
            float sx = v[0].x, sy = v[0].y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
            }
            v[0].x -= sx * (1.f / 32.f);
            v[0].y -= sy * (1.f / 32.f);
            float2 *dst = ex + r * Cfg::RROW + lane * Cfg::LROW;
            static_for<R1 / 2>([&](auto qi) {
                constexpr int q = decltype(qi)::value;
                const float2 y0 = q == 0 ? v[0] : cmul(v[2 * q], tw[2 * q]);
                const float2 y1 = cmul(v[2 * q + 1], tw[2 * q + 1]);
                *reinterpret_cast<float4 *>(dst + 2 * q) = make_float4(y0.x, y0.y, y1.x, y1.y);
            });
        }
    }
    __syncthreads();
    {
        const int r = threadIdx.x / R1, ka = threadIdx.x % R1;
        const float2 *src = ex + r * Cfg::RROW + ka;
        float2 u[32];
        static_for<32>([&](auto li) {
            constexpr int l = decltype(li)::value;
            u[brev<32>(l)] = src[l * Cfg::LROW];
        });
        fft_dit<32, +1>(u);
        // stage 03 shift + clip (rpv2.cu:137-148): column j' holds bin (j' + N/2) mod N, so
        // the zeroed columns N-1, N-2 are bins N/2-1 = (R1-1) + R1*15 and N/2-2.
        // stage 04 |.|^2 (rpv2.cu:150-157) and the row sum (rpv2.cu:171-197).
        float p = 0.f;
        static_for<32>([&](auto ki) {
            constexpr int kb = decltype(ki)::value;
            const float e = fmaf(u[kb].x, u[kb].x, u[kb].y * u[kb].y);
            if (kb == 15) {
                if (ka < R1 - 2) p += e;
            } else {
                p += e;
            }
        });
#pragma unroll
        for (int o = R1 / 2; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        // stages 05-08: circular convolution with the normalised taps, then /N; its row
        // sum is taps_sum * sum(p)
        p *= taps_sum;
        if (ka == 0) {
            p_row[r] = p;
            if (power) power[((size_t)sector * C + row_channel(r)) * half_m + row_gate(r)] = p;
        }
    }
    __syncthreads();
    // stages 09/10 (rpv2.cu:199-213)
    if (pair) {
        if (threadIdx.x < ROWS / 2) {
            const int g = blockIdx.x * (ROWS / 2) + threadIdx.x;
            const float p_hh = p_row[2 * threadIdx.x], p_vv = p_row[2 * threadIdx.x + 1];
            const float rg = (float)g * range_res;
            const float z = rg * rg * calib * p_hh;
            const float zdb = 10.f * log10f(z);
            const float zdr = 10.f * (log10f(p_hh) - log10f(p_vv));
            reinterpret_cast<float2 *>(out)[(size_t)sector * half_m + g] = make_float2(zdb, zdr);
        }
    } else if (C == 1) {
        if (threadIdx.x < ROWS) {
            const int g = ((int)blockIdx.x - pair_blocks) * ROWS + threadIdx.x;
            const float rg = (float)g * range_res;
            const float z = rg * rg * calib * p_row[threadIdx.x];
            reinterpret_cast<float2 *>(out)[(size_t)sector * half_m + g] = make_float2(10.f * log10f(z), 0.f);
        }
    }
}

// --------------------------------------------------------------------------------------
// wire records -> planar complex float.  One thread per record (12 bytes).
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    decode_wire_kernel(const uint32_t *__restrict__ wire, float2 *__restrict__ planar, size_t mn, int C,
                       size_t total_records)
{
    const size_t rec = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (rec >= total_records) return;
    const size_t sector = rec / mn, e = rec - sector * mn;
    const uint32_t *w = wire + rec * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        if (ch < C) {
            // bytes hiI loI hiQ loQ -> swap inside each half-word
            const uint32_t s = __byte_perm(__ldg(w + ch), 0u, 0x2301);
            const float fi = (float)(short)(s & 0xffffu);
            const float fq = (float)(short)(s >> 16);
            planar[(sector * C + ch) * mn + e] = make_float2(fi, fq);
        }
    }
}

// --------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------
bool fused_supported(int M, int N) { return M == 1024 && (N == 512 || N == 1024); }

using RangeA = RangeCfg<32, 32, 8>;

cudaError_t fused_setup()
{
    cudaError_t e;
    e = cudaFuncSetAttribute(range_fft_kernel<32, 32, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)RangeA::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(range_fft_kernel<32, 32, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)RangeA::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(doppler_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)DopplerCfg<16>::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(doppler_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)DopplerCfg<32>::SMEM);
    return e;
}

cudaError_t launch_decode_wire(const uint8_t *wire, float2 *planar, int M, int N, int C, int n_sectors,
                               cudaStream_t st)
{
    const size_t mn = (size_t)M * N, total = mn * n_sectors;
    if (total == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    decode_wire_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t *>(wire), planar, mn, C, total);
    return cudaGetLastError();
}

cudaError_t launch_range_fft(const float2 *iq, float2 *x2, const FusedTables &t, int M, int N, int C,
                             int n_sectors, cudaStream_t st)
{
    if (n_sectors == 0) return cudaSuccess;
    if (M != 1024 || N % 8) return cudaErrorInvalidValue;
    dim3 grid(N / 8, C * n_sectors);
    range_fft_kernel<32, 32, 8, false>
        <<<grid, RangeA::THREADS, RangeA::SMEM, st>>>(iq, x2, t.wrc_t, t.wd, t.tw_a, N);
    return cudaGetLastError();
}

cudaError_t launch_doppler(const float2 *x2, float *out, float *power, const FusedTables &t, int M, int N,
                           int C, int n_sectors, float range_res, float calib, float taps_sum,
                           cudaStream_t st)
{
    if (n_sectors == 0) return cudaSuccess;
    const int half_m = M / 2;
    if (N == 512) {
        using Cfg = DopplerCfg<16>;
        const int pair_blocks = C >= 2 ? half_m / (Cfg::ROWS / 2) : 0;
        const int single_blocks = (C & 1) ? half_m / Cfg::ROWS : 0;
        dim3 grid(pair_blocks + single_blocks, n_sectors);
        doppler_kernel<16><<<grid, 256, Cfg::SMEM, st>>>(x2, out, power, t.tw_b, half_m, C, pair_blocks,
                                                         range_res, calib, taps_sum);
    } else if (N == 1024) {
        using Cfg = DopplerCfg<32>;
        const int pair_blocks = C >= 2 ? half_m / (Cfg::ROWS / 2) : 0;
        const int single_blocks = (C & 1) ? half_m / Cfg::ROWS : 0;
        dim3 grid(pair_blocks + single_blocks, n_sectors);
        doppler_kernel<32><<<grid, 256, Cfg::SMEM, st>>>(x2, out, power, t.tw_b, half_m, C, pair_blocks,
                                                         range_res, calib, taps_sum);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace wrp
