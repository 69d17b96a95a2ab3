// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue/throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_scalar(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
    unsigned long long x[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; i++) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mixed(float* out, float a, float b) {  // packed + interleaved integer work
    unsigned long long x[8], aa, bb; int z[8];
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; i++) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); z[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb)); z[i] = (z[i] ^ it) + 3; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi + z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(d, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fl = 148.0 * 8 * 256 * 16 * ITERS * 2; printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(d, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0); k_mixed<<<148 * 8, 256>>>(d, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("packed+int  : %.3f ms  %.1f TFLOP/s (fp only)\n", ms, fl / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
