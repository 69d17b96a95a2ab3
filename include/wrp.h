/*
 * wrp.h — C ABI of libwrp, the B200-native per-sector weather-radar chain.
 *
 * This is the drop-in boundary for the hot path of rsatrioadi/weather-radar-processing:
 * IQ ingest -> Hamming window -> range FFT -> Doppler FFT (+fftshift, clip) -> |.|^2 ->
 * third FFT x moving-average coefficients -> inverse FFT -> power -> Z(dBZ)/ZDR.
 * The reference has no FFI of its own (every variant is a monolithic main()); each
 * entry point below names the reference function(s) it replaces, file:line relative
 * to the reference checkout.  Plain C types only; no CUDA or torch types appear in
 * any signature (streams and device pointers travel as void*).
 *
 * Conventions:
 *   - every call returns a wrp_status (0 = ok) and never exit()s — the reference's
 *     gpuErrchk -> exit(code) (rpv2.cu:21-27) becomes a status + wrp_last_error();
 *   - one handle per device per host thread; a handle is not thread-safe;
 *   - the handle owns the pinned ring, device buffers, tables and streams that the
 *     reference keeps in module globals (rpv2.cu:59-76) and frees them in wrp_destroy
 *     (rpv2.cu:685-722);
 *   - layouts are the reference's: input slot [sector][ch][i][j] complex float
 *     (rpv2.cu:379-381) or raw wire records (sector.cpp:52-62); result slot
 *     [sector][gate][2] = (ZdB, ZDR) floats (rpv2.cu:199-213, odim rpv2.cu:735).
 *   - there is NO CPU fallback: without a usable CUDA device wrp_create fails with
 *     WRP_ERR_CUDA.
 */
#ifndef WRP_H
#define WRP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WRP_VERSION 200 /* 2.0.0: wrp_config grew (doppler_form .. debug), streaming chain kernel, volume entry points */

typedef enum {
    WRP_OK = 0,
    WRP_ERR_INVALID = 1,     /* bad argument / NULL pointer / size out of range          */
    WRP_ERR_UNSUPPORTED = 2, /* M or N not a supported power of two, channels not 1..3    */
    WRP_ERR_CUDA = 3,        /* CUDA runtime error (text in wrp_last_error)               */
    WRP_ERR_NOMEM = 4,       /* host or device allocation failed                          */
    WRP_ERR_STATE = 5,       /* call sequence error (e.g. dump without a staged run)      */
    WRP_ERR_FULL = 6         /* wrp_submit: every ring slot is in flight, collect first   */
} wrp_status;

/* Input formats of one sector. */
typedef enum {
    /* planar complex float [ch][i][j], i<M rows (sweeps), j<N cols (samples): the
     * reference's pinned/device slot p_iq[j + i*N + ch*M*N] (rpv2.cu:379-381,
     * Dimension4::copy_at_depth dimension.cpp:19-21).  8*C*M*N bytes. */
    WRP_FMT_C64_PLANAR = 0,
    /* raw wire sector: M*N records of 12 bytes  hhI hhQ vvI vvQ vhI vhQ, big-endian
     * int16, record index i*N+j (sector.cpp:52-62, read_single.cc:145-172).  The
     * decode the reference does on one host thread (Sector::fromByteArray + the
     * restructuring loop rpv2.cu:369-383) happens on the GPU load path.  12*M*N bytes
     * regardless of n_channels (vh is skipped when n_channels == 2). */
    WRP_FMT_WIRE_I16BE = 1
} wrp_input_fmt;

typedef enum {
    /* ONE persistent kernel per batch, every intermediate in registers / shared memory
     * (chain_stream_kernel, csrc/wrp_stream.cu): a CTA walks the range tiles of one (sector,
     * channel) plane after another — window on load, radix-32 x radix-32 range FFT — and folds every
     * tile's rows k < M/2 straight into the per-gate sums that stages 03-08 reduce to in energy form
     * (Parseval: row energy minus the DC bin and the two clipped bins); ZdB/ZDR when a sector's hh
     * and vv planes are complete.  There is no range -> Doppler hand-off buffer and no CTA ever
     * waits for another.  Built for M = 1024 (planar or wire input) and M = 4096 (planar; wire goes
     * through a decode pre-pass), any power-of-two N in [64, 8192].  Wire records with three channels
     * run on chain_wire3_kernel (same file): a tile is 4 record columns x (hh, vv, vh), its raw
     * 48-byte rows arrive by TMA and are decoded in the first FFT pass (sector.cpp:52-62 on the GPU).
     * doppler_form = WRP_DOPPLER_FFT or chain_impl = WRP_CHAIN_QUEUE select the two-kind work queue
     * (chain_persistent_kernel: range tiles + Doppler blocks with an L2-resident hand-off ring,
     * M = 1024 / 4096, N = 512 / 1024), which can run the literal Doppler transform, shift, clip
     * and |.|^2.  Other shapes: WRP_ERR_UNSUPPORTED, use WRP_MODE_STAGED. */
    WRP_MODE_FUSED = 0,
    /* the reference's kernel cascade stage by stage (rpv2.cu:409-570), every stage
     * materialised in device memory so wrp_dump_stage can return 00iq..10zdr. */
    WRP_MODE_STAGED = 1
} wrp_mode;

/* How the fused kernels evaluate stages 03-08 (rpv2.cu:93-197). */
typedef enum {
    WRP_DOPPLER_ENERGY = 0, /* Parseval form: same products, no Doppler transform (default)      */
    WRP_DOPPLER_FFT = 1     /* literal mean removal, transform, shift, clip, |.|^2, row sum      */
} wrp_doppler_form;

/* Which fused kernel family carries the chain. */
typedef enum {
    WRP_CHAIN_AUTO = 0,   /* streaming kernel when the shape and doppler_form allow, else the queue */
    WRP_CHAIN_QUEUE = 1,  /* two-kind work queue with the L2-resident x2 ring (round-1 product path) */
    WRP_CHAIN_V1 = 2      /* two kernels per chunk (range_fft_kernel + doppler_kernel), M = 1024     */
} wrp_chain_impl;

/* Stage ids of the reference's dump files. */
typedef enum {
    WRP_STAGE_00_IQ = 0,   /* complex [M][N]   */
    WRP_STAGE_01_HAMM = 1, /* complex [M][N]   rpv2.cu:86-91    */
    WRP_STAGE_02_FFT1 = 2, /* complex [M][N]   rpv2.cu:426-428  */
    WRP_STAGE_03_FFT2 = 3, /* complex [M][N]   rpv2.cu:434-487  */
    WRP_STAGE_04_ABS = 4,  /* real    [M/2][N] rpv2.cu:150-157  */
    WRP_STAGE_05_FFT3 = 5, /* complex [M/2][N] rpv2.cu:510-512  */
    WRP_STAGE_06_MULT = 6, /* complex [M/2][N] rpv2.cu:159-163  */
    WRP_STAGE_07_CONV = 7, /* complex [M/2][N] rpv2.cu:526-528  */
    WRP_STAGE_08_POW = 8,  /* real    [M/2][N] rpv2.cu:165-169  */
    WRP_STAGE_09_ZDB = 9,  /* real    [M/2]    rpv2.cu:199-213  */
    WRP_STAGE_10_ZDR = 10, /* real    [M/2]    rpv2.cu:199-213  */
    WRP_STAGE_POWER = 11   /* real    [M/2]    row power P[i] (rpv2.cu:171-197), per channel */
} wrp_stage;

/* Replaces the compile-time constants of rpv2.cu:38-45 / radar_processor.h:17-25 and
 * the Dimension4 idim/odim set-up of rpv2.cu:734-736. */
typedef struct {
    int n_rows_M;      /* sweeps: range-FFT length  (rpv2.cu:40 n_sweeps = 1024)         */
    int n_cols_N;      /* samples: Doppler length   (rpv2.cu:41 n_samples = 512)         */
    int n_channels;    /* 3 = hh,vv,vh (rpv2.cu) ; 2 = hh,vv (read.cc)                   */
    int n_streams;     /* copy/compute streams and pinned ring depth (rpv2.cu argv[1])   */
    int ma_taps;       /* moving-average taps (rpv2.cu:45 ma_count = 7)                  */
    float range_res_m; /* rpv2.cu:43 k_range_resolution = 30                             */
    float calib;       /* rpv2.cu:44 k_calibration = 1941.05                             */
    int input_fmt;     /* wrp_input_fmt                                                  */
    int mode;          /* wrp_mode                                                       */
    int max_batch;     /* largest n_sectors of one process/submit call (ring slot size)  */
    int doppler_form;  /* wrp_doppler_form (fused mode)                                  */
    int chain_impl;    /* wrp_chain_impl (fused mode)                                    */
    int x2_lag;        /* queue kernel: Doppler blocks trail the range tiles by this many sectors (0 = default) */
    int x2_ring;       /* queue kernel: sector slots of the hand-off ring (0 = default)  */
    int evict_first;   /* queue kernel: stream the input through L2 evict-first (-1 = default, 0 off, 1 on) */
    int debug;         /* development switches: 16 = dependency counters of the queue kernel on stderr;
                        * 32 = planar input uses the one-channel wire kernel's work partition (bit-identical sums, tests);
                        * 128 = three-channel wire input keeps the one-channel-per-CTA kernel (A/B against chain_wire3_kernel);
                        * bits 8..17 = TMA L2 promotion of the planar tile loads in bytes (0 / 64 / 128 / 256; measured
                        * useless, default 0) */
} wrp_config;

typedef struct wrp_handle wrp_handle;

typedef struct {
    int version;
    int device;
    int sm_count;
    int l2_bytes;
    size_t input_bytes_per_sector;  /* in the configured input_fmt                       */
    size_t output_floats_per_sector; /* 2 * M/2                                          */
    size_t intermediate_bytes_per_sector; /* range->Doppler hand-off kept in L2 (0: streaming kernel) */
    int chunk_sectors;              /* sectors per launch (persistent kernel) or per kernel pair (v1) */
    int kernels_per_chunk;          /* launches of our kernels per chunk                 */
} wrp_info;

/* Accumulated device time per kernel family, measured with CUDA events on the
 * launching stream while profiling is enabled (bench.py's roofline figures). */
typedef struct {
    double ms_decode;     /* wire -> planar pre-pass (WRP_FMT_WIRE_I16BE only)            */
    double ms_range;      /* range-FFT kernel                                            */
    double ms_doppler;    /* Doppler/epilogue kernel                                     */
    double ms_staged;     /* staged cascade                                              */
    double ms_chain;      /* persistent fused-chain kernel (range + Doppler items)        */
    unsigned long long n_decode, n_range, n_doppler, n_staged, n_chain; /* launches measured */
    unsigned long long sectors;                                /* sectors processed       */
} wrp_profile;

/* Fill cfg with the reference's defaults (rpv2.cu:38-45): 1024 x 512 x 3, 7 taps,
 * 30 m, 1941.05, 3 streams, planar input, fused mode (streaming kernel, energy form), max_batch 8.
 * ALWAYS start from this call: fields added in later versions get their defaults here. */
void wrp_default_config(wrp_config *cfg);

/* Replaces generate_constants + prepare_arys + initialize_streams
 * (rpv2.cu:283-341; radar_processor.cu start() :48-57). */
int wrp_create(const wrp_config *cfg, int device, wrp_handle **out);

/* Replaces destroy_streams + destroy_arrays (rpv2.cu:685-722). NULL is a no-op. */
void wrp_destroy(wrp_handle *h);

/* Text of the last error on this handle (or of the last failed wrp_create when h
 * is NULL).  Never NULL. */
const char *wrp_last_error(const wrp_handle *h);

int wrp_get_info(const wrp_handle *h, wrp_info *info);

/* The init-time tables of generate_hamming_coefficients / generate_ma_coefficients
 * (rpv2.cu:222-281).  Any pointer may be NULL.  hamming: M*N floats; taps: ma_taps
 * floats; fft_ma: 2*N floats (re,im interleaved). */
int wrp_get_constants(const wrp_handle *h, float *hamming, float *taps, float *fft_ma);

/* HBM-resident batch: perform_stage_1/2/3 (rpv2.cu:409-570) for n_sectors sectors
 * whose input already sits in device memory in the configured format; writes
 * [n_sectors][M/2][2] floats to dev_out.  Asynchronous on cuda_stream (a
 * cudaStream_t passed as void*; NULL = the legacy default stream).  n_sectors may
 * be 0 (no-op) and may exceed max_batch (processed in chunks). */
int wrp_process_device(wrp_handle *h, const void *dev_iq, int n_sectors, float *dev_out,
                       void *cuda_stream);

/* Host-buffer batch: copy_matrix_to_device + stages + copy_result_to_host
 * (rpv2.cu:399-407, 409-570, 581-611) for n_sectors sectors in host memory, pipelined
 * over the handle's streams (H2D of chunk k+1 overlaps compute of chunk k and D2H of
 * chunk k-1).  host_iq should be pinned (wrp_alloc_pinned / cudaHostRegister);
 * pageable memory is staged through the handle's pinned ring.  Blocks until
 * host_out[n_sectors][M/2][2] is complete. */
int wrp_process_host(wrp_handle *h, const void *host_iq, int n_sectors, float *host_out);

/* Same pipeline, but the products stay in device memory: dev_out[n_sectors][M/2][2] on the handle's
 * device (no D2H).  What a multi-GPU caller uses before gathering the product volume over NVLink
 * (wrp_volume_process below, volume.py).  Blocks until dev_out is complete. */
int wrp_process_host_to_device(wrp_handle *h, const void *host_iq, int n_sectors, float *dev_out);

/* Streaming interface mirroring the reference's sector loop (rpv2.cu:665-683):
 * wrp_submit = read_matrix's hand-off + copy_matrix_to_device + the three stages for
 * up to max_batch sectors, tagged with their (sector, elevation) ids (advance(),
 * rpv2.cu:572-579); returns immediately.  wrp_collect = copy_result_to_host: blocks on
 * the oldest submission's event (no device-wide sync) and returns its products and
 * tags.  Ring depth is n_streams. */
int wrp_submit(wrp_handle *h, const void *host_iq, int n_sectors, const int *sector_ids,
               const int *elev_ids);
int wrp_collect(wrp_handle *h, float *out_zdb_zdr, int *sector_ids, int *elev_ids,
                int capacity_sectors, int *n_done);

/* Pinned host memory helpers (the reference's cudaMallocHost of p_iq, rpv2.cu:291). */
int wrp_alloc_pinned(size_t bytes, void **out);
int wrp_free_pinned(void *p);

/* Stage dump of sector `sector_in_batch` of the LAST batch processed in
 * WRP_MODE_STAGED (the reference's commented-out debug blocks, e.g. rpv2.cu:582-603).
 * host_out receives floats: complex stages interleaved (re,im).  *bytes is set to the
 * size written; with host_out == NULL only the size is returned. */
int wrp_dump_stage(wrp_handle *h, int sector_in_batch, int stage, int channel, void *host_out,
                   size_t *bytes);

/* Kernel launches of OUR kernels since creation (bench.py's gpu_launches). */
unsigned long long wrp_launch_count(const wrp_handle *h);

/* Name of the kernel that carries the chain for this handle's configuration
 * ("chain_stream_kernel", "chain_wire3_kernel", "chain_persistent_kernel", "range_fft_kernel", "staged cascade") —
 * what bench.py's roofline object and the ncu launch list refer to. */
const char *wrp_chain_kernel_name(const wrp_handle *h);

/* Diagnostic tap of the streaming kernel (tests): while dev_x2 is non-NULL every launch also
 * stores the range-FFT rows k < M/2 it folds — stage 02 as the PRODUCT kernel computes it — to
 * dev_x2[sector][channel][M/2][N] complex float (device memory owned by the caller, large enough
 * for the batches processed).  The reference's equivalent is the commented-out dump block inside
 * its running chain (rpv2.cu:582-603).  NULL switches the tap off.  WRP_ERR_STATE if the handle
 * does not run the streaming kernel. */
int wrp_set_stage02_tap(wrp_handle *h, void *dev_x2);

/* The fused gather (the reference collects every sector's products into one result[]
 * volume, rpv2.cu:607, 736).  While n_mirrors > 0, wrp_process_device and wrp_process_host_to_device
 * store every (ZdB, ZDR) pair not only at dev_out[i] but also at mirrors[m][i] for every m — the
 * same float index relative to the mirror as relative to dev_out of that call.  A mirror is any
 * device-accessible address: normally the slice "this device's units" of the product volume that
 * lives on ANOTHER device of the box (peer access enabled, or a CUDA-IPC mapping of another
 * process's buffer), so that the streaming kernel's epilogue writes the volume over NVLink itself
 * and no collective or copy follows it.  The stores are complete when the stream work of the call
 * has completed (the usual CUDA rule for peer writes); a consumer on the other device still needs
 * that one event / barrier.  Kernels other than the streaming ones copy dev_out to the mirrors on
 * the same stream instead.  The pointer array is copied; n_mirrors = 0 switches it off.
 * wrp_process_host, wrp_submit/wrp_collect and the staged dumps ignore mirrors. */
#define WRP_MAX_PRODUCT_MIRRORS 8
int wrp_set_product_mirrors(wrp_handle *h, float *const *mirrors, int n_mirrors);

/* Per-kernel CUDA-event timing. enable: 1 start accumulating / 0 stop. */
int wrp_profile_enable(wrp_handle *h, int enable);
int wrp_profile_read(wrp_handle *h, wrp_profile *out, int reset);

/* Product serialisation of send_results (rpv2.cu:620-663, floats.c:3-36):
 * header [sector BE16][elev BE16] (with_elev != 0, rpv2: 4 + 4*gates bytes) or
 * [sector BE16] (gpu_1fp_streamcasc.cu:709-716: 2 + 4*gates bytes), followed by
 * `gates` big-endian floats.  zdb_zdr is one result slot [gates][2].  Returns the
 * packet size in bytes, or a negative wrp_status. Host-only, needs no handle. */
int wrp_pack_products(const float *zdb_zdr, int gates, int sector, int elev, int with_elev,
                      uint8_t *zdb_packet, uint8_t *zdr_packet);

int wrp_version(void);

/* ---- volume scan over several devices of one box (BASELINE config 4) -----------------
 * The reference walks (elevation, sector) units one at a time on one GPU (advance(), rpv2.cu:572-579)
 * and stores unit k = elevation * n_sectors + sector at result[sitdim(0, 0, sector, elevation)]
 * (rpv2.cu:607, 736).  Units are independent, so a volume is cut into contiguous unit blocks
 * [ceil(U g / G), ceil(U (g+1) / G)) — one per device, each keeping a contiguous slice of the
 * reference's result array — with NO data-path collective.  wrp_volume_process runs one host thread
 * per device (own handle, pinned ring and streams), leaves every device's products in its memory,
 * gathers the slices into the volume buffer on devices[0] with peer copies (NVLink where the
 * devices are peers) and returns the volume with one D2H copy.  A device may be listed more than
 * once (several shards on one GPU). */
typedef struct wrp_volume wrp_volume;
int wrp_volume_create(const wrp_config *cfg, const int *devices, int n_devices, int n_sectors, int n_elevations,
                      wrp_volume **out);
/* units [*first_unit, *first_unit + *n_units) belong to shard `shard` (0 <= shard < n_devices) */
int wrp_volume_shard(const wrp_volume *v, int shard, int *first_unit, int *n_units);
/* host_iq: n_elevations * n_sectors sectors in unit order, configured input_fmt (pinned recommended);
 * host_volume: [n_elevations][n_sectors][M/2][2] floats = the reference's result[] (rpv2.cu:736). */
int wrp_volume_process(wrp_volume *v, const void *host_iq, float *host_volume);
const char *wrp_volume_last_error(const wrp_volume *v);
void wrp_volume_destroy(wrp_volume *v);

#ifdef __cplusplus
}
#endif
#endif /* WRP_H */
