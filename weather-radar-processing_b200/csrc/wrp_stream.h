// wrp_stream.h — parameters and host entry points of the streaming chain kernel (wrp_stream.cu):
// the fused chain WITHOUT a range -> Doppler hand-off.  Not part of the C ABI.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace wrp {

struct FusedTables;

struct StreamParams {
    // tables (device)
    const float *wrc_t;     // [32 b][32 a] wr(32 a + b) * c          (M = 1024: window of the first butterfly stage)
    const float2 *tw_a;     // [32 b][32 ka] exp(-2 pi i b ka / 1024)  (inter-pass twiddles of the 32 x 32 transform)
    const float *wd;        // [N] Doppler window wd(j)
    const float *wr4;       // M = 4096: wr(i) * c, natural order
    const float2 *tw4;      // M = 4096: exp(-2 pi i r / 4096), r < 1024 (radix-4 pre-pass)
    const float4 *tile_tw;  // [N / T] (cos, sin) of -2 pi (T t + (T-1)/2) m / N for m = 1 and m = 2: a tile's factor of the
                            // clipped bins (its first column's phase times the centring factor of the column pairing)
    // data
    const void *in;         // planar [S][C][M][N] float2, or wire records [S][M][N] x 12 B
    float *out;             // [S][M/2][2]  (ZdB, ZDR)
    float *mirror[8];       // the fused gather (wrp_set_product_mirrors): every product is also stored at mirror[i][same index];
    int n_mirrors;          // peer-mapped memory of the other devices of the box (NVLink) or plain device memory
    float *power;           // [S][C][M/2]  row powers (also the hh/vv exchange between the CTAs that finish the planes)
    float2 *x2_tap;         // optional debug tap: range-FFT rows k < M/2, [S][C][M/2][N]; NULL in production
    float *scratch;         // [grid][2][7][M/2] partial sums of planes shared between CTAs
    int *plane_cnt;         // [S * C] parts of a shared plane that have arrived
    int *sector_cnt;        // [S] finished (hh, vv) planes of a sector
    int S, C, N, NT;        // sectors, channels, Doppler length, tiles per plane (N / T)
    int half_m;             // M / 2
    int chan_groups;        // 1: CTAs split the [plane][tile] space; C: CTA x works on channel x % C of the
                            // [sector][tile] space, so the C CTAs that read the same wire records run side by side
    int zero;               // always 0; makes a TMA issue data-dependent on a loaded value (tma_load_2d)
    float n_float, range_res, calib, taps_sum;
    float2 wcol[2][8];      // [m][c], c < T/2: (-1)^c (cos, sin)(2 pi m ((T-1)/2 - c) / N): the pair (c, T-1-c)'s factors of
                            // clipped bin N/2 - m (fold_row in wrp_stream.cu)
};

bool stream_supported(int M, int N, int wire);
// occupancy-derived grid (CTAs that are certainly co-resident are not required: no CTA ever waits for another)
cudaError_t stream_setup(int M, int wire, int sm_count, int *max_grid);
size_t stream_scratch_floats(int M, int max_grid);
// channel_groups != 0 forces the wire path's work partition (CTA x -> channel x % C) on planar input too,
// so that both formats associate their sums identically (bit-exactness tests of the decode)
// tmap: planar input only — 2-D tensor map over the batch viewed as [S*C*M rows][N] 8-byte elements with
// box {T columns, 8192 / (8 T) rows} (stream_encode_tensor_map); wire input passes a zeroed map
cudaError_t launch_stream(StreamParams p, int M, int wire, int max_grid, int sm_count, int channel_groups,
                          const CUtensorMap &tmap, cudaStream_t st);
// Encodes the tensor map of one launch (host-side, no CUDA call besides the driver's encoder).
// encode_fn = cuTensorMapEncodeTiled obtained through cudaGetDriverEntryPoint (libwrp does not link libcuda).
// l2_promotion_bytes: 0 / 64 / 128 / 256 — the tensor map's L2 promotion: a row segment narrower than that pulls the
// whole aligned block into L2, so the following tiles of the plane (the same CTA, a few microseconds later) hit
bool stream_encode_tensor_map(void *encode_fn, CUtensorMap *out, const void *base, int M, int N, long long planes,
                              int l2_promotion_bytes);
const char *stream_kernel_name();

// chain_wire3_kernel: wire records, M = 1024, three channels in one CTA (12-column tiles, raw rows by TMA).
// tile_tw / wcol must be built for 4-column tiles.
cudaError_t wire3_setup(int sm_count, int *max_grid);
size_t wire3_scratch_floats(int max_grid);
bool wire3_encode_tensor_map(void *encode_fn, CUtensorMap *out, const void *base, int N, long long sectors);
cudaError_t launch_wire3(StreamParams p, int max_grid, int sm_count, const CUtensorMap &tmap, cudaStream_t st);

} // namespace wrp
