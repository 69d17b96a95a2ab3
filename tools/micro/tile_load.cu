// tile_load.cu — micro-benchmark: how fast can one SM pull pulse x gate tiles out of a row-major
// [rows][N] matrix into shared memory, by load mechanism and tile geometry?  (VERDICT r1 item 3:
// "128-byte-row tile staging, with TMA re-examined on that geometry".)
//
// Every CTA is a persistent loop over tiles with NBUF shared-memory buffers in flight; a tile is
// "consumed" by a token read of every 512th byte (so nothing can be elided), i.e. the figure is the
// pure load path: issue + L2/DRAM + the shared-memory write side.
//
//   mode 0  cp.async 16 B per lane (LDGSTS), what chain_stream_kernel does for planar input
//   mode 1  TMA tiled boxes  {COLS x 8 B inner, 256 rows} x (ROWS / 256)     (cp.async.bulk.tensor.2d)
//   mode 2  cp.async 4 B per lane out of 12-byte wire records (one channel), chain_stream_kernel's wire path
//   mode 3  TMA boxes over the RAW wire records {COLS x 12 B inner, 256 rows}: all three channels of the
//           columns land as [rows][COLS x 12 B] (the channel is picked apart by the shared-memory reads).
//           (TMA cannot gather one channel: elementStrides[0] is ignored without interleave — tried, the
//           engine then writes the whole 96-byte rows past the 32-byte-row buffer — and an inner box of one
//           4-byte element violates the 16-byte minimum.)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tile_load tile_load.cu
// run:   ./tile_load            (prints one line per configuration)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                       \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) {                                                                    \
            printf("%s: %s\n", #x, cudaGetErrorString(e_));                                         \
            exit(1);                                                                                \
        }                                                                                           \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

struct Params {
    const uint8_t *in;
    int n_cols;     // N (elements per matrix row: complex floats, or wire records)
    int rows;       // rows per tile (1024 or 4096)
    int cols;       // columns per tile
    int n_tiles;    // total tiles
    int tiles_per_plane;
    int mode, nbuf;
    int assign;     // 0: CTA x takes tiles x, x + G, ...; 1: a contiguous run per CTA; g >= 2: groups of g CTAs share a
                    // contiguous run and interleave its tiles (the g CTAs read g adjacent row segments at the same time)
    unsigned long long *sink;
};

// tile t covers rows [plane * rows, +rows) x columns [ct * cols, +cols)
__global__ void __launch_bounds__(256) load_kernel(const Params p, const __grid_constant__ CUtensorMap map)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int elem = p.mode == 2 ? 4 : (p.mode == 3 ? 12 : 8); // bytes landed per (row, column)
    const int tile_bytes = p.rows * p.cols * elem;
    const int pitch = p.cols * elem;
    if (tid == 0)
        for (int i = 0; i < p.nbuf; ++i) mbar_init(&bar[i], (p.mode == 0 || p.mode == 2) ? blockDim.x : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    auto issue = [&](int t, int buf) {
        const int plane = t / p.tiles_per_plane, ct = t - plane * p.tiles_per_plane;
        uint8_t *dst = smem + (size_t)buf * tile_bytes;
        if (p.mode == 0) {
            const int cpr = pitch / 16;                       // 16-byte chunks per row
            const int rows_per_pass = blockDim.x / cpr;
            const uint8_t *src = p.in + ((size_t)plane * p.rows + tid / cpr) * ((size_t)p.n_cols * 8) + (size_t)ct * pitch + (tid % cpr) * 16;
            for (int r = 0; r < p.rows; r += rows_per_pass)
                cp_async16(dst + (size_t)(r + tid / cpr) * pitch + (tid % cpr) * 16, src + (size_t)r * p.n_cols * 8);
            cp_async_arrive(&bar[buf]);
        } else if (p.mode == 2) {
            const int rows_per_pass = blockDim.x / p.cols;
            const uint8_t *src = p.in + ((size_t)plane * p.rows + tid / p.cols) * ((size_t)p.n_cols * 12) + ((size_t)ct * p.cols + tid % p.cols) * 12;
            for (int r = 0; r < p.rows; r += rows_per_pass)
                cp_async4(dst + (size_t)(r + tid / p.cols) * pitch + (tid % p.cols) * 4, src + (size_t)r * p.n_cols * 12);
            cp_async_arrive(&bar[buf]);
        } else if (tid == 0) {
            mbar_expect_tx(&bar[buf], tile_bytes);
            for (int r = 0; r < p.rows; r += 256) {
                if (p.mode == 1) tma_load_2d(dst + (size_t)r * pitch, &map, ct * p.cols, plane * p.rows + r, &bar[buf]);
                else tma_load_2d(dst + (size_t)r * pitch, &map, ct * p.cols * 3, plane * p.rows + r, &bar[buf]); // uint32 elements
            }
        }
    };

    unsigned long long acc = 0;
    int n_mine = 0;
    const int G = gridDim.x, x = blockIdx.x;
    auto tile_of = [&](int k) -> int {
        if (p.assign == 0) return x + k * G;
        if (p.assign == 1) return (int)((long long)p.n_tiles * x / G) + k;
        const int g = p.assign, grp = x / g, j = x - grp * g;
        return (int)((long long)p.n_tiles * (grp * g) / G) + k * g + j;
    };
    if (p.assign == 0) {
        for (int t = x; t < p.n_tiles; t += G) ++n_mine;
    } else if (p.assign == 1) {
        n_mine = (int)((long long)p.n_tiles * (x + 1) / G) - (int)((long long)p.n_tiles * x / G);
    } else {
        const int g = p.assign, grp = x / g, j = x - grp * g;
        const int lo = (int)((long long)p.n_tiles * (grp * g) / G), hi = (int)((long long)p.n_tiles * min((grp + 1) * g, G) / G);
        n_mine = (hi - lo - j + g - 1) / g;
    }
    for (int k = 0; k < p.nbuf - 1 && k < n_mine; ++k) issue(tile_of(k), k);
    for (int k = 0; k < n_mine; ++k) {
        const int buf = k % p.nbuf;
        if (k + p.nbuf - 1 < n_mine) {
            __syncthreads(); // the buffer being refilled was consumed by everybody one iteration ago
            issue(tile_of(k + p.nbuf - 1), (k + p.nbuf - 1) % p.nbuf);
        }
        mbar_wait(&bar[buf], (k / p.nbuf) & 1);
        const uint8_t *src = smem + (size_t)buf * tile_bytes;
        for (int o = tid * 512; o < tile_bytes; o += blockDim.x * 512) acc += *reinterpret_cast<const uint32_t *>(src + o);
    }
    if (acc == 0x1234567ull) p.sink[0] = acc + lane + warp;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int S = argc > 1 ? atoi(argv[1]) : 96; // planes of 1024 x 512 (4 MiB each as complex float)
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    const size_t bytes = (size_t)S * 3 * 1024 * 512 * 8; // also enough for the 4096 x 1024 and wire views
    uint8_t *d = nullptr;
    unsigned long long *sink = nullptr;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 1, bytes));
    CK(cudaMalloc(&sink, 8));
    CK(cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));

    struct Cfg {
        const char *name;
        int mode, rows, cols, n_cols, nbuf, ctas_per_sm, threads;
        int assign = 0;
    };
    const std::vector<Cfg> cfgs = {
        {"cp.async16   8 col x 1024 rows (64 B rows), 1 buf, 2 CTA/SM", 0, 1024, 8, 512, 1, 2, 256},
        {"cp.async16   8 col x 1024 rows (64 B rows), 1 buf, 3 CTA/SM", 0, 1024, 8, 512, 1, 3, 256},
        {"cp.async16   8 col x 1024 rows (64 B rows), 2 buf, 1 CTA/SM", 0, 1024, 8, 512, 2, 1, 256},
        {"cp.async16  16 col x 1024 rows (128 B rows), 1 buf, 1 CTA/SM", 0, 1024, 16, 512, 1, 1, 256},
        {"cp.async16   4 col x 4096 rows (32 B rows), 1 buf, 1 CTA/SM", 0, 4096, 4, 1024, 1, 1, 256},
        {"TMA tiled    8 col x 1024 rows (64 B rows), 1 buf, 2 CTA/SM", 1, 1024, 8, 512, 1, 2, 256},
        {"TMA tiled    8 col x 1024 rows (64 B rows), 1 buf, 3 CTA/SM", 1, 1024, 8, 512, 1, 3, 256},
        {"TMA tiled    8 col x 1024 rows (64 B rows), 2 buf, 1 CTA/SM", 1, 1024, 8, 512, 2, 1, 256},
        {"TMA tiled   16 col x 1024 rows (128 B rows), 1 buf, 1 CTA/SM", 1, 1024, 16, 512, 1, 1, 256},
        {"TMA tiled   32 col x 512 rows (256 B rows), 1 buf, 1 CTA/SM", 1, 512, 32, 512, 1, 1, 256},
        {"TMA tiled    4 col x 4096 rows (32 B rows), 1 buf, 1 CTA/SM", 1, 4096, 4, 1024, 1, 1, 256},
        {"TMA tiled    4 col x 4096 rows (32 B rows), contiguous run per CTA", 1, 4096, 4, 1024, 1, 1, 256, 1},
        {"TMA tiled    4 col x 4096 rows (32 B rows), groups of 2 CTAs interleave", 1, 4096, 4, 1024, 1, 1, 256, 2},
        {"TMA tiled    4 col x 4096 rows (32 B rows), groups of 4 CTAs interleave", 1, 4096, 4, 1024, 1, 1, 256, 4},
        {"TMA tiled    4 col x 4096 rows (32 B rows), groups of 8 CTAs interleave", 1, 4096, 4, 1024, 1, 1, 256, 8},
        {"TMA tiled    8 col x 1024 rows (64 B rows), 1 buf, 2 CTA/SM, contiguous run", 1, 1024, 8, 512, 1, 2, 256, 1},
        {"cp.async4   wire, 8 col x 1024 rows, 1 buf, 2 CTA/SM", 2, 1024, 8, 512, 1, 2, 256},
        {"cp.async4   wire, 8 col x 1024 rows, 2 buf, 2 CTA/SM", 2, 1024, 8, 512, 2, 2, 256},
        {"TMA raw wire rows, 8 col x 1024 rows (96 B rows), 1 buf, 2 CTA/SM", 3, 1024, 8, 512, 1, 2, 256},
        {"TMA raw wire rows, 4 col x 1024 rows (48 B rows), 2 buf, 2 CTA/SM", 3, 1024, 4, 512, 2, 2, 256},
        {"TMA raw wire rows, 4 col x 1024 rows (48 B rows), 1 buf, 4 CTA/SM", 3, 1024, 4, 512, 1, 4, 256},
    };
    printf("%-66s %10s %10s %12s\n", "configuration", "us/tile/SM", "GB/s", "rows/us/SM");
    for (const Cfg &c : cfgs) {
        const bool wire = c.mode >= 2;
        const int elem = c.mode == 2 ? 4 : (c.mode == 3 ? 12 : 8);
        const size_t row_bytes = (size_t)c.n_cols * (wire ? 12 : 8);
        const size_t plane_bytes = row_bytes * c.rows;
        const int planes = (int)(bytes / plane_bytes);
        Params p{};
        p.in = d;
        p.n_cols = c.n_cols;
        p.rows = c.rows;
        p.cols = c.cols;
        p.tiles_per_plane = c.n_cols / c.cols;
        p.n_tiles = planes * p.tiles_per_plane;
        p.mode = c.mode;
        p.nbuf = c.nbuf;
        p.assign = c.assign;
        p.sink = sink;
        CUtensorMap map{};
        if (c.mode == 1 || c.mode == 3) {
            cuuint64_t dims[2], strides[1];
            cuuint32_t box[2], estr[2];
            CUtensorMapDataType dt;
            if (c.mode == 1) { // elements = 8-byte complex floats
                dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
                dims[0] = c.n_cols, dims[1] = (cuuint64_t)planes * c.rows;
                strides[0] = row_bytes;
                box[0] = c.cols, box[1] = 256;
                estr[0] = 1, estr[1] = 1;
            } else { // elements = 32-bit (I, Q) pairs, three per record
                dt = CU_TENSOR_MAP_DATA_TYPE_UINT32;
                dims[0] = (cuuint64_t)c.n_cols * 3, dims[1] = (cuuint64_t)planes * c.rows;
                strides[0] = row_bytes;
                box[0] = c.cols * 3, box[1] = 256;
                estr[0] = 1, estr[1] = 1;
            }
            CUresult r = encode(&map, dt, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                printf("%-66s tensor map rejected (CUresult %d)\n", c.name, (int)r);
                continue;
            }
        }
        const int tile_bytes = c.rows * c.cols * elem;
        const int smem = tile_bytes * c.nbuf;
        const int grid = sms * c.ctas_per_sm;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int it = 0; it < 2; ++it) load_kernel<<<grid, c.threads, smem>>>(p, map);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 5;
        for (int it = 0; it < reps; ++it) load_kernel<<<grid, c.threads, smem>>>(p, map);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        const double useful = (double)p.n_tiles * tile_bytes;
        const double us_tile_sm = ms * 1e3 / ((double)p.n_tiles / sms);
        printf("%-66s %10.3f %10.0f %12.1f\n", c.name, us_tile_sm, useful / (ms * 1e-3) / 1e9, c.rows / us_tile_sm);
    }
    return 0;
}
