// wrp_chain_params.h — kernel parameters and control-word layout shared by the persistent chain
// kernels (wrp_persistent.cu: queue of range tiles and Doppler blocks; wrp_unified.cu: one item =
// one range tile + eight Doppler rows).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace wrp {

// ---- parameters ----------------------------------------------------------------------------
struct PersistParams {
    const float *wrc_t;
    const float *wd;
    const float2 *tw_a;
    const float2 *tw_b;
    const float *wr4;   // M = 4096 only: wr(i)*c, natural order [4096]
    const float2 *tw4;  // M = 4096 only: exp(-2*pi*i*r/4096), [1024]
    const float2 *iq; // input [S][C][M][N]
    float2 *x2;       // ring [ring][C][M/2][N]
    float *out;       // [S][M/2][2]
    float *power;     // optional [S][C][M/2]
    int *ctrl;        // [0] work counter; [1..2] debug; a_done at CTRL_A; b_done at CTRL_A + smax
    int S, C, N, half_m;
    int ring, lag;
    int tiles_a, blocks_b, pair_blocks;
    int n1, n2, n3, b3_first; // queue regions (see decode_item)
    int total_items;
    int smax;
    int evict_first; // stream the input through L2 with an evict-first policy (WRP_EVICT_FIRST=0 turns it off)
    int discard; // drop consumed ring rows from L2 with discard.global.L2 (WRP_DISCARD=1 turns it on)
    int debug; // WRP_DEBUG development switches
    int zero;  // always 0 (see opaque_smem_addr)
    float range_res, calib, taps_sum;
};
constexpr int CTRL_A = 32;

// Which kernel carries a fused batch of this shape under the current environment switches
// (WRP_CHAIN=queue, WRP_DOPPLER=fft, WRP_TILE_COLS=4, WRP_DISCARD=1 and an L2 access-policy window
// all select the two-kind queue of wrp_persistent.cu).  The one place that decides: the launch,
// the lag / ring defaults and wrp_chain_kernel_name all ask here.
bool chain_uses_unified_kernel(int M, int N, size_t l2_window_bytes);

// wrp_unified.cu: M = 1024, N = 512 only.  p.tiles_a, p.pair_blocks are ignored (recomputed).
bool unified_supported(int M, int N);
cudaError_t unified_setup();
cudaError_t launch_unified(PersistParams p, int sm_count, cudaStream_t st);

} // namespace wrp
