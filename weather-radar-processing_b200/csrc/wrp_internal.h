// wrp_internal.h — handle layout and kernel-launch prototypes shared by the
// translation units of libwrp.  Not part of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/wrp.h"

namespace wrp {

// Host-side tables (wrp_tables.cpp) — the reference's generate_constants (rpv2.cu:222-287).
struct HostTables {
    std::vector<float> ham;       // [M][N]  wr(i)*wd(j)*c, rpv2.cu:245-249 (double math, float store)
    std::vector<float> wr_c;      // [M]     wr(i)*c
    std::vector<float> wd;        // [N]     wd(j)
    std::vector<float> taps;      // [ma_taps]
    std::vector<float> fft_ma;    // [N][2]  forward N-point DFT of the zero-padded taps
    std::vector<float> tw_m;      // [M][2]  exp(-2*pi*i*t/M)   (forward, range)
    std::vector<float> tw_n;      // [N][2]  exp(+2*pi*i*t/N)   (inverse, Doppler)
    double c = 0.0;               // window normalisation constant (negative)
    float taps_sum = 1.f;         // sum of the float taps (== 1 up to rounding)
};
void build_host_tables(int M, int N, int ma_taps, HostTables &t);

// Device tables of the fused kernels.
struct FusedTables {
    float *wrc_t = nullptr;  // [R2a][R1a]   wr_c[R2a*a + b] stored as [b][a] (range pass-1 window)
    float *wd = nullptr;     // [N]
    float2 *tw_a = nullptr;  // [R2a][R1a]   exp(-2*pi*i*b*ka/M)    range inter-pass twiddles
    float2 *tw_b = nullptr;  // [32][R1b]    exp(+2*pi*i*l*ka/N)    Doppler inter-pass twiddles
    float *wr4 = nullptr;    // M = 4096: wr_c in natural order (radix-4 pre-pass window)
    float2 *tw4 = nullptr;   // M = 4096: [1024] exp(-2*pi*i*r/4096) (pre-pass twiddle)
    float4 *tile_tw = nullptr; // streaming kernel: [N/T] per-tile factors of the two clipped Doppler bins
};

struct StagedBuffers {
    // all [batch][C][...] ; complex stages float2
    float2 *s00 = nullptr, *s01 = nullptr, *s02 = nullptr, *s03 = nullptr; // [C][M][N]
    float *s04 = nullptr, *s08 = nullptr;                                  // [C][M/2][N]
    float2 *s05 = nullptr, *s06 = nullptr, *s07 = nullptr;                 // [C][M/2][N]
    float2 *rowsum = nullptr;                                              // [C][M] complex row sums (d_tmp, rpv2.cu:299)
    float *power = nullptr;                                                // [C][M/2]
    float *result = nullptr;                                               // [batch][M/2][2] products of the last run
    float *ham = nullptr;                                                  // [M][N]
    float2 *fft_ma = nullptr;                                              // [N]
    float2 *tw_m = nullptr, *tw_n_fwd = nullptr, *tw_n_inv = nullptr;      // twiddle tables
    int batch_capacity = 0;
    int last_batch = 0; // sectors of the last staged run (for wrp_dump_stage)
};

struct RingSlot {
    void *pinned_in = nullptr;   // max_batch sectors of input
    float *pinned_out = nullptr; // max_batch result slots
    void *dev_in = nullptr;
    float *dev_out = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;     // products of this slot are in pinned_out
    cudaEvent_t h2d_done = nullptr; // input of this slot is in dev_in
    int n_sectors = 0;          // >0 while in flight
    std::vector<int> sector_ids, elev_ids;
};

} // namespace wrp

struct wrp_handle {
    wrp_config cfg{};
    int device = 0;
    int sm_count = 0;
    int l2_bytes = 0;
    std::string err;

    wrp::HostTables host;
    wrp::FusedTables fused;
    wrp::StagedBuffers staged;

    // which kernel family carries a fused batch (decided once, at wrp_create)
    enum ChainKind { CHAIN_STREAM, CHAIN_QUEUE, CHAIN_V1, CHAIN_STAGED };
    ChainKind chain = CHAIN_STAGED;
    bool decode_prepass = false; // wire input is decoded into `decoded` before the chain kernel
    // streaming kernel (wrp_stream.cu): no hand-off; partial sums of planes cut by the work partition,
    // per-plane / per-sector arrival counters, row powers [smax][C][M/2]
    int stream_max_grid = 0;
    bool wire3 = false; // wire input, M = 1024, 3 channels: chain_wire3_kernel (12-column tiles, raw rows by TMA)
    int l2_promotion = 0; // tensor-map L2 promotion of the tile loads, bytes
    void *tma_encode = nullptr; // cuTensorMapEncodeTiled (driver entry point; libwrp does not link libcuda)
    float *stream_scratch = nullptr;
    int *stream_cnt = nullptr;
    float2 wcol[2][8] = {};
    float2 *x2_tap = nullptr; // wrp_set_stage02_tap (caller-owned)
    float *mirrors[WRP_MAX_PRODUCT_MIRRORS] = {}; // wrp_set_product_mirrors (caller-owned, usually peer-mapped)
    int n_mirrors = 0;
    // queue / v1 kernels: range -> Doppler hand-off, [ring or chunk][C][M/2][N] float2, L2-resident
    float2 *x2 = nullptr;
    float2 *decoded = nullptr; // [chunk][C][M][N] planar scratch for wire input (decode pre-pass)
    float *power = nullptr;    // [smax or chunk][C][M/2]
    int chunk = 1;             // sectors per launch
    int x2_ring = 8; // queue kernel: sector slots of the x2 hand-off ring (50 MB, L2-resident)
    int x2_lag = 4;  // queue kernel: Doppler blocks of sector t are queued after the range tiles of sector t + lag
    int *ctrl = nullptr;
    int smax = 1024; // sectors per persistent launch

    cudaStream_t compute_stream = nullptr; // all kernels of the host path run here (owns the scratch)
    std::vector<wrp::RingSlot> ring;
    int ring_head = 0; // next slot to submit into
    int ring_tail = 0; // oldest in-flight slot
    int ring_inflight = 0;

    unsigned long long launches = 0;
    bool profiling = false;
    wrp_profile prof{};
    struct PendingEvent {
        cudaEvent_t a, b;
        int kind;
    };
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;
};

namespace wrp {

// Fused path (wrp_fused.cu).  Each returns cudaGetLastError() after the launch.
bool fused_supported(int M, int N);
cudaError_t fused_setup(); // cudaFuncSetAttribute for the big-smem kernels
cudaError_t launch_decode_wire(const uint8_t *wire, float2 *planar, int M, int N, int C, int n_sectors,
                               cudaStream_t st);
cudaError_t launch_range_fft(const float2 *iq, float2 *x2, const FusedTables &t, int M, int N, int C,
                             int n_sectors, cudaStream_t st);
cudaError_t launch_doppler(const float2 *x2, float *out, float *power, const FusedTables &t, int M, int N,
                           int C, int n_sectors, float range_res, float calib, float taps_sum,
                           cudaStream_t st);

// Persistent fused chain (wrp_persistent.cu): one launch per batch of <= smax sectors.
bool persistent_supported(int M, int N);
cudaError_t persistent_setup();
int persistent_ctrl_ints(int smax);
cudaError_t launch_persistent(const float2 *iq, float *out, float *power, float2 *x2_ring, int ring, int lag,
                              int *ctrl,
                              int smax, const FusedTables &t, int M, int N, int C, int n_sectors, float range_res,
                              float calib, float taps_sum, int sm_count, bool doppler_fft, int evict_first, int debug,
                              cudaStream_t st);

void persistent_debug_counters(const int *ctrl, int *not_ready_a, int *not_ready_b);

// Staged path (wrp_staged.cu): the reference cascade, one stage per kernel.
// Returns the number of kernels launched through *launches.
cudaError_t staged_setup();
cudaError_t run_staged(wrp_handle *h, const void *dev_in, int n_sectors, float *dev_out, cudaStream_t st,
                       unsigned long long *launches);

} // namespace wrp
