// wrp_staged.cu — WRP_MODE_STAGED: the reference's kernel cascade, one kernel per stage,
// every stage materialised so wrp_dump_stage can return the 00iq..10zdr dumps.
//
// This is the dump/verification path, generic over power-of-two M, N (<= 8192); speed is
// not its purpose (the fused path in wrp_fused.cu is the product path).  Stage semantics
// follow rpv2.cu:409-570 one to one; the three cuFFT plans (rpv2.cu:318-341) are replaced
// by one shared-memory Stockham radix-2 line-FFT kernel.
#include "wrp_internal.h"

namespace wrp {

// ---- generic line FFT ----------------------------------------------------------------
// One CTA per line of L points: element q of line (outer, inner) sits at
// base = outer*outer_pitch + inner*inner_pitch, then + q*stride.  tw[t] = exp(sign 2 pi i t/L).
__global__ void k_fft_lines(float2 *data, const float2 *__restrict__ tw, int L, size_t stride, int n_inner,
                            size_t inner_pitch, size_t outer_pitch)
{
    extern __shared__ __align__(16) float2 sm[];
    float2 *x = sm, *y = sm + L;
    const size_t line = blockIdx.x;
    const size_t outer = line / n_inner, inner = line % n_inner;
    float2 *base = data + outer * outer_pitch + inner * inner_pitch;
    for (int q = threadIdx.x; q < L; q += blockDim.x) x[q] = base[(size_t)q * stride];
    __syncthreads();
    for (int l = L / 2, m = 1; l >= 1; l >>= 1, m <<= 1) {
        for (int idx = threadIdx.x; idx < L / 2; idx += blockDim.x) {
            const int j = idx / m, k = idx - j * m;
            const float2 c0 = x[k + j * m], c1 = x[k + j * m + l * m];
            const float2 w = tw[j * m];
            const float dx = c0.x - c1.x, dy = c0.y - c1.y;
            y[k + 2 * j * m] = make_float2(c0.x + c1.x, c0.y + c1.y);
            y[k + 2 * j * m + m] = make_float2(dx * w.x - dy * w.y, dx * w.y + dy * w.x);
        }
        __syncthreads();
        float2 *t = x;
        x = y;
        y = t;
    }
    for (int q = threadIdx.x; q < L; q += blockDim.x) base[(size_t)q * stride] = x[q];
}

// ---- element-wise stages -----------------------------------------------------------
// stage 01: __apply_hamming (rpv2.cu:86-91); idx % mn is radar_processor.cu:20-25's form
__global__ void k_apply_hamming(const float2 *__restrict__ in, float2 *__restrict__ out,
                                const float *__restrict__ ham, size_t mn, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float b = ham[i % mn];
    const float2 a = in[i];
    out[i] = make_float2(b * a.x, b * a.y);
}

// __sum_v4 (rpv2.cu:93-121): complex sum of each row; one CTA (256 threads) per row
__global__ void k_row_sum_complex(const float2 *__restrict__ in, float2 *__restrict__ sums, int N)
{
    __shared__ float2 sd[256];
    const size_t row = blockIdx.x;
    float2 acc = make_float2(0.f, 0.f);
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float2 v = in[row * N + j];
        acc.x += v.x;
        acc.y += v.y;
    }
    sd[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sd[threadIdx.x].x += sd[threadIdx.x + s].x;
            sd[threadIdx.x].y += sd[threadIdx.x + s].y;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[row] = sd[0];
}

// __avgconj (rpv2.cu:123-130): x = conj(x - sum/N)
__global__ void k_avgconj(const float2 *__restrict__ in, float2 *__restrict__ out,
                          const float2 *__restrict__ sums, int N, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float2 s = sums[i / N];
    const float avgx = s.x / (float)N, avgy = s.y / (float)N;
    const float2 v = in[i];
    out[i] = make_float2(v.x - avgx, (v.y - avgy) * -1.f);
}

// __conjugate + __shift + __clip_v2 (rpv2.cu:132-148) == __conjshift (gpu_1fp_uni.cu:106-115)
// one thread per pair (j, j + N/2)
__global__ void k_conjshift_clip(float2 *data, int N, size_t total_pairs)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_pairs) return;
    const int half = N / 2;
    const size_t row = i / half;
    const int j = (int)(i - row * half);
    float2 *r = data + row * N;
    const float2 lo = r[j], hi = r[j + half];
    float2 new_lo = make_float2(hi.x, -hi.y), new_hi = make_float2(lo.x, -lo.y);
    if (j + half >= N - 2) new_hi = make_float2(0.f, 0.f); // columns N-1, N-2
    if (j >= N - 2) new_lo = make_float2(0.f, 0.f);        // only when N <= 4
    r[j] = new_lo;
    r[j + half] = new_hi;
}

// __abssqr (rpv2.cu:150-157) on rows < M/2: s04 real, s05 seeded with (p, 0)
__global__ void k_abssqr(const float2 *__restrict__ s03, float *__restrict__ s04, float2 *__restrict__ s05,
                         int M, int N, size_t total_half)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_half) return;
    const size_t hmn = (size_t)(M / 2) * N;
    const size_t plane = i / hmn, e = i - plane * hmn;
    const float2 v = s03[plane * (size_t)M * N + e];
    const float p = v.x * v.x + v.y * v.y;
    s04[i] = p;
    s05[i] = make_float2(p, 0.f);
}

// __apply_ma (rpv2.cu:159-163)
__global__ void k_apply_ma(const float2 *__restrict__ s05, float2 *__restrict__ s06,
                           const float2 *__restrict__ fft_ma, int N, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float2 a = s05[i], b = fft_ma[i % N];
    s06[i] = make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// __scale_real (rpv2.cu:165-169)
__global__ void k_scale_real(const float2 *__restrict__ s07, float *__restrict__ s08, int N, size_t total)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    s08[i] = s07[i].x / (float)N;
}

// __sum_inplace_v4 (rpv2.cu:171-197): P[row] = sum_j s08[row][j]
__global__ void k_row_sum_real(const float *__restrict__ in, float *__restrict__ sums, int N)
{
    __shared__ float sd[256];
    const size_t row = blockIdx.x;
    float acc = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) acc += in[row * N + j];
    sd[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sd[threadIdx.x] += sd[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[row] = sd[0];
}

// __calcresult_v2 (rpv2.cu:199-213); power is [sector][C][M/2]
__global__ void k_calcresult(const float *__restrict__ power, float *__restrict__ out, int half_m, int C,
                             float range_res, float calib, size_t total_gates)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_gates) return;
    const size_t sector = i / half_m;
    const int g = (int)(i - sector * half_m);
    const float p_hh = power[(sector * C + 0) * half_m + g];
    const float rg = (float)g * range_res;
    const float z = rg * rg * calib * p_hh;
    const float zdb = 10.f * log10f(z);
    float zdr = 0.f;
    if (C >= 2) {
        const float p_vv = power[(sector * C + 1) * half_m + g];
        zdr = 10.f * (log10f(p_hh) - log10f(p_vv));
    }
    out[2 * i] = zdb;
    out[2 * i + 1] = zdr;
}

static inline unsigned nblk(size_t total, int threads = 256) { return (unsigned)((total + threads - 1) / threads); }

#define WRP_LAUNCH_CHECK()                     \
    do {                                       \
        cudaError_t e__ = cudaGetLastError();  \
        if (e__ != cudaSuccess) return e__;    \
        ++n;                                   \
    } while (0)

cudaError_t run_staged(wrp_handle *h, const void *dev_in, int S, float *dev_out, cudaStream_t st,
                       unsigned long long *launches)
{
    const int M = h->cfg.n_rows_M, N = h->cfg.n_cols_N, C = h->cfg.n_channels;
    StagedBuffers &b = h->staged;
    const size_t mn = (size_t)M * N, hmn = (size_t)(M / 2) * N;
    const size_t planes = (size_t)S * C;
    const size_t total = planes * mn, total_half = planes * hmn;
    unsigned long long n = 0;
    cudaError_t e;
    if (S == 0) {
        *launches = 0;
        return cudaSuccess;
    }

    // stage 00: ingest
    if (h->cfg.input_fmt == WRP_FMT_WIRE_I16BE) {
        e = launch_decode_wire((const uint8_t *)dev_in, b.s00, M, N, C, S, st);
        if (e != cudaSuccess) return e;
        ++n;
    } else {
        e = cudaMemcpyAsync(b.s00, dev_in, total * sizeof(float2), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
    }
    // stage 01
    k_apply_hamming<<<nblk(total), 256, 0, st>>>(b.s00, b.s01, b.ham, mn, total);
    WRP_LAUNCH_CHECK();
    // stage 02: range FFT along i (stride N) for every column of every plane, rpv2.cu:318-333
    e = cudaMemcpyAsync(b.s02, b.s01, total * sizeof(float2), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    {
        const int threads = M / 2 < 512 ? M / 2 : 512;
        k_fft_lines<<<(unsigned)(planes * N), threads, 2 * (size_t)M * sizeof(float2), st>>>(
            b.s02, b.tw_m, M, (size_t)N, N, 1, mn);
        WRP_LAUNCH_CHECK();
    }
    // stage 03: row mean, subtract + conjugate, forward FFT, conjugate + shift + clip
    k_row_sum_complex<<<(unsigned)(planes * M), 256, 0, st>>>(b.s02, b.rowsum, N);
    WRP_LAUNCH_CHECK();
    k_avgconj<<<nblk(total), 256, 0, st>>>(b.s02, b.s03, b.rowsum, N, total);
    WRP_LAUNCH_CHECK();
    {
        const int threads = N / 2 < 512 ? N / 2 : 512;
        k_fft_lines<<<(unsigned)(planes * M), threads, 2 * (size_t)N * sizeof(float2), st>>>(
            b.s03, b.tw_n_fwd, N, 1, 1, 0, (size_t)N);
        WRP_LAUNCH_CHECK();
    }
    k_conjshift_clip<<<nblk(total / 2), 256, 0, st>>>(b.s03, N, total / 2);
    WRP_LAUNCH_CHECK();
    // stage 04 (+ seed of 05)
    k_abssqr<<<nblk(total_half), 256, 0, st>>>(b.s03, b.s04, b.s05, M, N, total_half);
    WRP_LAUNCH_CHECK();
    // stage 05: forward FFT of the power rows
    {
        const int threads = N / 2 < 512 ? N / 2 : 512;
        k_fft_lines<<<(unsigned)(planes * (M / 2)), threads, 2 * (size_t)N * sizeof(float2), st>>>(
            b.s05, b.tw_n_fwd, N, 1, 1, 0, (size_t)N);
        WRP_LAUNCH_CHECK();
    }
    // stage 06
    k_apply_ma<<<nblk(total_half), 256, 0, st>>>(b.s05, b.s06, b.fft_ma, N, total_half);
    WRP_LAUNCH_CHECK();
    // stage 07: un-normalised inverse FFT
    e = cudaMemcpyAsync(b.s07, b.s06, total_half * sizeof(float2), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    {
        const int threads = N / 2 < 512 ? N / 2 : 512;
        k_fft_lines<<<(unsigned)(planes * (M / 2)), threads, 2 * (size_t)N * sizeof(float2), st>>>(
            b.s07, b.tw_n_inv, N, 1, 1, 0, (size_t)N);
        WRP_LAUNCH_CHECK();
    }
    // stage 08
    k_scale_real<<<nblk(total_half), 256, 0, st>>>(b.s07, b.s08, N, total_half);
    WRP_LAUNCH_CHECK();
    // row power, stages 09/10
    k_row_sum_real<<<(unsigned)(planes * (M / 2)), 256, 0, st>>>(b.s08, b.power, N);
    WRP_LAUNCH_CHECK();
    k_calcresult<<<nblk((size_t)S * (M / 2)), 256, 0, st>>>(b.power, b.result, M / 2, C, h->cfg.range_res_m,
                                                           h->cfg.calib, (size_t)S * (M / 2));
    WRP_LAUNCH_CHECK();
    e = cudaMemcpyAsync(dev_out, b.result, (size_t)S * M * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    b.last_batch = S;
    *launches = n;
    return cudaSuccess;
}

cudaError_t staged_setup()
{
    return cudaFuncSetAttribute(k_fft_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 8);
}

} // namespace wrp
