// wrp_chain_params.h — kernel parameters and control-word layout of the two-kind work-queue kernel
// (wrp_persistent.cu: queue of range tiles and Doppler blocks).
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
    int evict_first; // stream the input through L2 with an evict-first policy (wrp_config.evict_first)
    int discard; // drop consumed ring rows from L2 with discard.global.L2 (off: measured 2 % slower)
    int debug; // wrp_config.debug development switches
    int zero;  // always 0 (see opaque_smem_addr)
    float range_res, calib, taps_sum;
};
constexpr int CTRL_A = 32;

} // namespace wrp
