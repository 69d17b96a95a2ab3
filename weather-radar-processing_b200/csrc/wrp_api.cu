// wrp_api.cu — the C ABI of include/wrp.h: handle life cycle, the HBM-resident batch
// path, the pinned multi-stream host path (submit/collect ring), stage dumps, profiling.
//
// What it replaces in the reference: the module globals and their set-up/tear-down
// (rpv2.cu:59-76, 283-341, 685-722), the sector loop (do_process, rpv2.cu:665-683) and the
// copy helpers (copy_matrix_to_device :399-407, copy_result_to_host :581-611).
// Differences by design: errors are returned, not exit()ed; completion is tracked with
// per-slot events instead of cudaDeviceSynchronize after every launch group
// (rpv2.cu:422-491); scratch is owned by one compute stream so the reference's shared
// d_tmp race cannot occur.
#include "wrp_internal.h"
#include "wrp_chain_params.h"
#include "wrp_stream.h"

#include <nvtx3/nvToolsExt.h> // header-only: ranges cost nothing unless a profiler is attached

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

static std::string g_create_error = "";

#define CK(h, call)                                                                           \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            return e__ == cudaErrorMemoryAllocation ? WRP_ERR_NOMEM : WRP_ERR_CUDA;           \
        }                                                                                     \
    } while (0)

// NVTX range per C-ABI call on the hot path — the reference's tick()/tock() hook points
// (gpu_1fp.cu:173-185, 279-286) as profiler ranges
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static int fail(wrp_handle *h, int code, const std::string &msg)
{
    h->err = msg;
    return code;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static size_t input_bytes_per_sector(const wrp_config &c)
{
    const size_t mn = (size_t)c.n_rows_M * c.n_cols_N;
    return c.input_fmt == WRP_FMT_WIRE_I16BE ? mn * 12 : mn * 8 * c.n_channels;
}

template <typename T> static cudaError_t upload(T **dst, const void *src, size_t bytes)
{
    cudaError_t e = cudaMalloc((void **)dst, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
}

extern "C" {

int wrp_version(void) { return WRP_VERSION; }

void wrp_default_config(wrp_config *cfg)
{
    if (!cfg) return;
    cfg->n_rows_M = 1024;
    cfg->n_cols_N = 512;
    cfg->n_channels = 3;
    cfg->n_streams = 3;
    cfg->ma_taps = 7;
    cfg->range_res_m = 30.f;
    cfg->calib = 1941.05f;
    cfg->input_fmt = WRP_FMT_C64_PLANAR;
    cfg->mode = WRP_MODE_FUSED;
    cfg->max_batch = 8;
    cfg->doppler_form = WRP_DOPPLER_ENERGY;
    cfg->chain_impl = WRP_CHAIN_AUTO;
    cfg->x2_lag = 0;
    cfg->x2_ring = 0;
    cfg->evict_first = -1;
    cfg->debug = 0;
}

const char *wrp_last_error(const wrp_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static void free_all(wrp_handle *h)
{
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    auto F = [](void *p) {
        if (p) cudaFree(p);
    };
    F(h->fused.wrc_t);
    F(h->fused.wd);
    F(h->fused.tw_a);
    F(h->fused.tw_b);
    F(h->fused.wr4);
    F(h->fused.tw4);
    F(h->fused.tile_tw);
    wrp::StagedBuffers &b = h->staged;
    F(b.s00), F(b.s01), F(b.s02), F(b.s03), F(b.s04), F(b.s05), F(b.s06), F(b.s07), F(b.s08);
    F(b.rowsum), F(b.power), F(b.result), F(b.ham), F(b.fft_ma), F(b.tw_m), F(b.tw_n_fwd), F(b.tw_n_inv);
    F(h->x2), F(h->decoded), F(h->power), F(h->ctrl), F(h->stream_scratch), F(h->stream_cnt);
    for (auto &s : h->ring) {
        if (s.pinned_in) cudaFreeHost(s.pinned_in);
        if (s.pinned_out) cudaFreeHost(s.pinned_out);
        F(s.dev_in), F(s.dev_out);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    }
    if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
    for (auto &p : h->pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto e : h->event_pool) cudaEventDestroy(e);
}

static int create_impl(wrp_handle *h)
{
    const wrp_config &c = h->cfg;
    const int M = c.n_rows_M, N = c.n_cols_N, C = c.n_channels;
    CK(h, cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CK(h, cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    h->l2_bytes = prop.l2CacheSize;

    wrp::build_host_tables(M, N, c.ma_taps, h->host);
    const wrp::HostTables &t = h->host;

    const size_t mn = (size_t)M * N, hmn = (size_t)(M / 2) * N;
    if (c.mode == WRP_MODE_FUSED) {
        // transposed window [b][a] = wr_c[32 a + b]; inter-pass twiddles [b][ka] (range) and
        // [l][ka] (Doppler)
        // (M = 4096: the kernel's radix-4 pre-pass leaves 1024-point sub-transforms, whose 32 x 32
        // twiddles are every 4th entry of the 4096-point table; the window is applied in the pre-pass)
        const int QM = M / 1024, R1a = 32, R2a = 32, R1b = N / 32;
        std::vector<float> wrc_t(1024);
        std::vector<float> tw_a(2 * 1024), tw_b(2 * (size_t)N);
        for (int b = 0; b < R2a; b++)
            for (int a = 0; a < R1a; a++) {
                wrc_t[(size_t)b * R1a + a] = t.wr_c[R2a * a + b];
                const int q = ((b * a) % 1024) * QM; // ka = a
                tw_a[2 * ((size_t)b * R1a + a)] = t.tw_m[2 * q];
                tw_a[2 * ((size_t)b * R1a + a) + 1] = t.tw_m[2 * q + 1];
            }
        if (QM == 4) {
            std::vector<float> tw4(t.tw_m.begin(), t.tw_m.begin() + 2 * 1024); // exp(-2 pi i r / 4096), r < 1024
            CK(h, upload(&h->fused.wr4, t.wr_c.data(), (size_t)M * 4));
            CK(h, upload(&h->fused.tw4, tw4.data(), tw4.size() * 4));
        }
        for (int l = 0; l < 32; l++)
            for (int ka = 0; ka < R1b; ka++) {
                const int q = (l * ka) % N;
                tw_b[2 * ((size_t)l * R1b + ka)] = t.tw_n[2 * q];
                tw_b[2 * ((size_t)l * R1b + ka) + 1] = t.tw_n[2 * q + 1];
            }
        CK(h, upload(&h->fused.wrc_t, wrc_t.data(), wrc_t.size() * 4));
        CK(h, upload(&h->fused.wd, t.wd.data(), t.wd.size() * 4));
        CK(h, upload(&h->fused.tw_a, tw_a.data(), tw_a.size() * 4));
        CK(h, upload(&h->fused.tw_b, tw_b.data(), tw_b.size() * 4));
        CK(h, wrp::fused_setup());
        const size_t inter = (size_t)C * hmn * sizeof(float2);
        const bool wire = c.input_fmt == WRP_FMT_WIRE_I16BE;
        const bool want_fft = c.doppler_form == WRP_DOPPLER_FFT;
        // kernel family: decided here, once (no environment switches at launch time)
        if (c.chain_impl == WRP_CHAIN_V1) {
            if (!wrp::fused_supported(M, N) || want_fft) {
                h->err = "wrp_create: the two-kernel form (WRP_CHAIN_V1) supports M = 1024 and the literal Doppler transform only via WRP_CHAIN_QUEUE";
                return WRP_ERR_UNSUPPORTED;
            }
            h->chain = wrp_handle::CHAIN_V1;
        } else if (c.chain_impl == WRP_CHAIN_AUTO && !want_fft && wrp::stream_supported(M, N, 0)) {
            h->chain = wrp_handle::CHAIN_STREAM;
        } else if (wrp::persistent_supported(M, N)) {
            h->chain = wrp_handle::CHAIN_QUEUE;
        } else {
            h->err = "wrp_create: no fused kernel for this shape / doppler_form / chain_impl; use WRP_MODE_STAGED";
            return WRP_ERR_UNSUPPORTED;
        }
        if (h->chain == wrp_handle::CHAIN_STREAM) {
            // the streaming kernel decodes wire records on its load path (M = 1024); M = 4096 wire
            // input goes through the decode pre-pass
            const int wire_direct = wire && wrp::stream_supported(M, N, 1);
            h->decode_prepass = wire && !wire_direct;
            h->smax = 1024;
            h->chunk = h->smax;
            h->wire3 = wire_direct && C == 3 && !(c.debug & 128); // debug 128: keep the one-channel-per-CTA wire kernel (A/B)
            if (h->wire3) CK(h, wrp::wire3_setup(h->sm_count, &h->stream_max_grid));
            else CK(h, wrp::stream_setup(M, wire_direct, h->sm_count, &h->stream_max_grid));
            // experiment knob (debug = bytes << 8): tensor-map L2 promotion of the tile loads.  Measured on B200:
            // 128 B no change, 64 B / 256 B slower, on both shapes — left off.
            h->l2_promotion = (c.debug >> 8) & 0x3ff;
            if (!wire_direct || h->wire3) { // planar tiles and raw wire rows are fetched by TMA: the tensor map of a launch is encoded on the host
                cudaDriverEntryPointQueryResult q;
                CK(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->tma_encode, cudaEnableDefault, &q));
                if (!h->tma_encode) return fail(h, WRP_ERR_CUDA, "wrp_create: the driver does not export cuTensorMapEncodeTiled");
            }
            const int T = (M == 4096 || h->wire3) ? 4 : 8, NT = N / T; // columns per tile
            std::vector<float> ttw(4 * (size_t)NT);
            const double kTwoPi = 6.283185307179586476925286766559;
            for (int tt = 0; tt < NT; tt++)
                for (int m = 1; m <= 2; m++) {
                    // e^{-2 pi i m (T tt + (T-1)/2) / N}: first column of the tile, times the centring factor of the pairing
                    const double a = -kTwoPi * m * ((double)(((long long)T * tt) % N) + 0.5 * (T - 1)) / N;
                    ttw[4 * (size_t)tt + 2 * (m - 1)] = (float)std::cos(a);
                    ttw[4 * (size_t)tt + 2 * (m - 1) + 1] = (float)std::sin(a);
                }
            CK(h, upload(&h->fused.tile_tw, ttw.data(), ttw.size() * 4));
            for (int m = 1; m <= 2; m++)
                for (int cc = 0; cc < 8; cc++) { // pair (cc, T-1-cc), cc < T/2 (the rest is unused)
                    const double phi = kTwoPi * m * (0.5 * (T - 1) - cc) / N, sg = (cc & 1) ? -1.0 : 1.0;
                    h->wcol[m - 1][cc] = make_float2((float)(sg * std::cos(phi)), (float)(sg * std::sin(phi)));
                }
            const size_t scratch_floats =
                h->wire3 ? wrp::wire3_scratch_floats(h->stream_max_grid) : wrp::stream_scratch_floats(M, h->stream_max_grid);
            CK(h, cudaMalloc((void **)&h->stream_scratch, scratch_floats * sizeof(float)));
            CK(h, cudaMalloc((void **)&h->stream_cnt, sizeof(int) * (size_t)h->smax * (C + 1)));
            CK(h, cudaMalloc((void **)&h->power, (size_t)h->smax * C * (M / 2) * sizeof(float)));
        } else if (h->chain == wrp_handle::CHAIN_QUEUE) {
            h->decode_prepass = wire;
            // x2 hand-off = ring of sector slots that stays in L2 (ring * C*(M/2)*N*8 bytes); lag 4 / ring 8
            // measured best for 1024 x 512
            if (M == 4096) { // 24-48 MiB of hand-off per sector: keep as few sectors in flight as the queue allows
                h->x2_lag = 1;
                h->x2_ring = 3;
            }
            if (c.x2_ring > 0) h->x2_ring = c.x2_ring;
            if (c.x2_lag > 0) h->x2_lag = c.x2_lag;
            if (h->x2_ring < h->x2_lag + 2) h->x2_ring = h->x2_lag + 2; // a tile may only wait for earlier queue items
            if (h->x2_ring > 64) h->x2_ring = 64;
            h->smax = 1024;
            h->chunk = h->smax;
            CK(h, wrp::persistent_setup());
            CK(h, cudaMalloc((void **)&h->x2, inter * h->x2_ring));
            CK(h, cudaMalloc((void **)&h->ctrl, sizeof(int) * wrp::persistent_ctrl_ints(h->smax)));
        } else {
            h->decode_prepass = wire;
            // chunk: sectors per kernel pair, sized so the range->Doppler hand-off
            // (C*(M/2)*N*8 bytes per sector) stays resident in L2 between the two kernels
            int chunk = (int)((size_t)h->l2_bytes / 3 / inter);
            if (chunk < 1) chunk = 1;
            if (chunk > 4096) chunk = 4096;
            h->chunk = chunk;
            CK(h, cudaMalloc((void **)&h->x2, inter * chunk));
            CK(h, cudaMalloc((void **)&h->power, (size_t)chunk * C * (M / 2) * sizeof(float)));
        }
        if (h->decode_prepass) {
            // planar scratch of the decode pre-pass, allocated once: it bounds the launch size, so
            // wrp_process_device never allocates or synchronises
            int dchunk = c.max_batch > 8 ? c.max_batch : 8;
            if (dchunk > h->chunk) dchunk = h->chunk;
            h->chunk = dchunk;
            CK(h, cudaMalloc((void **)&h->decoded, (size_t)h->chunk * C * mn * sizeof(float2)));
        }
    } else {
        wrp::StagedBuffers &b = h->staged;
        b.batch_capacity = c.max_batch;
        const size_t P = (size_t)c.max_batch * C;
        CK(h, cudaMalloc((void **)&b.s00, P * mn * 8));
        CK(h, cudaMalloc((void **)&b.s01, P * mn * 8));
        CK(h, cudaMalloc((void **)&b.s02, P * mn * 8));
        CK(h, cudaMalloc((void **)&b.s03, P * mn * 8));
        CK(h, cudaMalloc((void **)&b.s04, P * hmn * 4));
        CK(h, cudaMalloc((void **)&b.s05, P * hmn * 8));
        CK(h, cudaMalloc((void **)&b.s06, P * hmn * 8));
        CK(h, cudaMalloc((void **)&b.s07, P * hmn * 8));
        CK(h, cudaMalloc((void **)&b.s08, P * hmn * 4));
        CK(h, cudaMalloc((void **)&b.rowsum, P * M * 8));
        CK(h, cudaMalloc((void **)&b.power, P * (M / 2) * 4));
        CK(h, cudaMalloc((void **)&b.result, (size_t)c.max_batch * M * 4));
        CK(h, upload(&b.ham, t.ham.data(), mn * 4));
        CK(h, upload(&b.fft_ma, t.fft_ma.data(), (size_t)N * 8));
        CK(h, upload(&b.tw_m, t.tw_m.data(), (size_t)M * 8));
        CK(h, upload(&b.tw_n_inv, t.tw_n.data(), (size_t)N * 8));
        std::vector<float> fwd(t.tw_n);
        for (int q = 0; q < N; q++) fwd[2 * q + 1] = -fwd[2 * q + 1];
        CK(h, upload(&b.tw_n_fwd, fwd.data(), (size_t)N * 8));
        CK(h, wrp::staged_setup());
        h->chunk = c.max_batch;
    }
    CK(h, cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
    h->ring.resize(c.n_streams);
    return WRP_OK;
}

int wrp_create(const wrp_config *cfg, int device, wrp_handle **out)
{
    if (!cfg || !out) {
        g_create_error = "wrp_create: NULL argument";
        return WRP_ERR_INVALID;
    }
    *out = nullptr;
    const wrp_config &c = *cfg;
    if (c.n_channels < 1 || c.n_channels > 3 || c.n_streams < 1 || c.n_streams > 64 || c.ma_taps < 1 ||
        c.ma_taps > 63 || c.max_batch < 1 || c.max_batch > 4096 ||
        (c.input_fmt != WRP_FMT_C64_PLANAR && c.input_fmt != WRP_FMT_WIRE_I16BE) ||
        (c.mode != WRP_MODE_FUSED && c.mode != WRP_MODE_STAGED) ||
        (c.doppler_form != WRP_DOPPLER_ENERGY && c.doppler_form != WRP_DOPPLER_FFT) || c.chain_impl < WRP_CHAIN_AUTO ||
        c.chain_impl > WRP_CHAIN_V1 || c.x2_lag < 0 || c.x2_ring < 0 || c.x2_lag > 32 || c.x2_ring > 64) {
        g_create_error = "wrp_create: configuration field out of range";
        return WRP_ERR_INVALID;
    }
    if (!is_pow2(c.n_rows_M) || !is_pow2(c.n_cols_N) || c.n_rows_M < 4 || c.n_cols_N < 4 ||
        c.n_rows_M > 8192 || c.n_cols_N > 8192 || c.ma_taps > c.n_cols_N) {
        g_create_error = "wrp_create: M and N must be powers of two in [4, 8192]";
        return WRP_ERR_UNSUPPORTED;
    }
    if (c.mode == WRP_MODE_FUSED && !wrp::fused_supported(c.n_rows_M, c.n_cols_N) &&
        !wrp::persistent_supported(c.n_rows_M, c.n_cols_N) && !wrp::stream_supported(c.n_rows_M, c.n_cols_N, 0)) {
        g_create_error = "wrp_create: fused mode supports M=1024 or 4096 with a power-of-two N in [64, 8192]; use WRP_MODE_STAGED";
        return WRP_ERR_UNSUPPORTED;
    }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || device < 0 || device >= n_dev) {
        g_create_error = std::string("wrp_create: no usable CUDA device (") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                         "); libwrp has no CPU fallback";
        cudaGetLastError();
        return e != cudaSuccess ? WRP_ERR_CUDA : WRP_ERR_INVALID;
    }
    wrp_handle *h = new (std::nothrow) wrp_handle();
    if (!h) {
        g_create_error = "wrp_create: out of host memory";
        return WRP_ERR_NOMEM;
    }
    h->cfg = c;
    h->device = device;
    const int rc = create_impl(h);
    if (rc != WRP_OK) {
        g_create_error = h->err;
        free_all(h);
        delete h;
        return rc;
    }
    *out = h;
    return WRP_OK;
}

void wrp_destroy(wrp_handle *h)
{
    if (!h) return;
    free_all(h);
    delete h;
}

int wrp_get_info(const wrp_handle *h, wrp_info *info)
{
    if (!h || !info) return WRP_ERR_INVALID;
    const wrp_config &c = h->cfg;
    info->version = WRP_VERSION;
    info->device = h->device;
    info->sm_count = h->sm_count;
    info->l2_bytes = h->l2_bytes;
    info->input_bytes_per_sector = input_bytes_per_sector(c);
    info->output_floats_per_sector = (size_t)c.n_rows_M;
    info->intermediate_bytes_per_sector =
        h->chain == wrp_handle::CHAIN_STREAM ? 0 : (size_t)c.n_channels * (c.n_rows_M / 2) * c.n_cols_N * 8;
    info->chunk_sectors = h->chunk;
    const int pre = h->decode_prepass ? 1 : 0;
    info->kernels_per_chunk = c.mode == WRP_MODE_FUSED ? (h->chain == wrp_handle::CHAIN_V1 ? 2 : 1) + pre
                                                       : 14 + (c.input_fmt == WRP_FMT_WIRE_I16BE ? 1 : 0);
    return WRP_OK;
}

int wrp_get_constants(const wrp_handle *h, float *hamming, float *taps, float *fft_ma)
{
    if (!h) return WRP_ERR_INVALID;
    const wrp::HostTables &t = h->host;
    if (hamming) memcpy(hamming, t.ham.data(), t.ham.size() * sizeof(float));
    if (taps) memcpy(taps, t.taps.data(), t.taps.size() * sizeof(float));
    if (fft_ma) memcpy(fft_ma, t.fft_ma.data(), t.fft_ma.size() * sizeof(float));
    return WRP_OK;
}

unsigned long long wrp_launch_count(const wrp_handle *h) { return h ? h->launches : 0ull; }

const char *wrp_chain_kernel_name(const wrp_handle *h)
{
    if (!h) return "";
    switch (h->chain) {
    case wrp_handle::CHAIN_STREAM: return h->wire3 ? "chain_wire3_kernel" : wrp::stream_kernel_name();
    case wrp_handle::CHAIN_QUEUE: return "chain_persistent_kernel";
    case wrp_handle::CHAIN_V1: return "range_fft_kernel";
    default: return "staged cascade";
    }
}

int wrp_set_stage02_tap(wrp_handle *h, void *dev_x2)
{
    if (!h) return WRP_ERR_INVALID;
    if (h->chain != wrp_handle::CHAIN_STREAM)
        return fail(h, WRP_ERR_STATE, "wrp_set_stage02_tap: the handle does not run the streaming kernel");
    h->x2_tap = (float2 *)dev_x2;
    return WRP_OK;
}

int wrp_set_product_mirrors(wrp_handle *h, float *const *mirrors, int n_mirrors)
{
    if (!h) return WRP_ERR_INVALID;
    if (n_mirrors < 0 || n_mirrors > WRP_MAX_PRODUCT_MIRRORS || (n_mirrors > 0 && !mirrors))
        return fail(h, WRP_ERR_INVALID, "wrp_set_product_mirrors: n_mirrors must be in [0, WRP_MAX_PRODUCT_MIRRORS]");
    for (int m = 0; m < n_mirrors; m++) {
        if (!mirrors[m] || ((uintptr_t)mirrors[m] & 7)) // the kernels store (ZdB, ZDR) pairs
            return fail(h, WRP_ERR_INVALID, "wrp_set_product_mirrors: a mirror is NULL or not 8-byte aligned");
        h->mirrors[m] = mirrors[m];
    }
    h->n_mirrors = n_mirrors;
    return WRP_OK;
}

// ---- profiling ------------------------------------------------------------------------
static cudaEvent_t get_event(wrp_handle *h)
{
    if (!h->event_pool.empty()) {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    wrp_handle *h;
    cudaStream_t st;
    int kind;
    cudaEvent_t a = nullptr;
    ProfScope(wrp_handle *h_, cudaStream_t st_, int kind_) : h(h_), st(st_), kind(kind_)
    {
        if (h->profiling) {
            a = get_event(h);
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope()
    {
        if (a) {
            cudaEvent_t b = get_event(h);
            cudaEventRecord(b, st);
            h->pending.push_back({a, b, kind});
        }
    }
};

int wrp_profile_enable(wrp_handle *h, int enable)
{
    if (!h) return WRP_ERR_INVALID;
    h->profiling = enable != 0;
    return WRP_OK;
}

int wrp_profile_read(wrp_handle *h, wrp_profile *out, int reset)
{
    if (!h || !out) return WRP_ERR_INVALID;
    CK(h, cudaSetDevice(h->device));
    for (auto &p : h->pending) {
        CK(h, cudaEventSynchronize(p.b));
        float ms = 0.f;
        CK(h, cudaEventElapsedTime(&ms, p.a, p.b));
        switch (p.kind) {
        case 0: h->prof.ms_decode += ms, h->prof.n_decode++; break;
        case 1: h->prof.ms_range += ms, h->prof.n_range++; break;
        case 2: h->prof.ms_doppler += ms, h->prof.n_doppler++; break;
        case 4: h->prof.ms_chain += ms, h->prof.n_chain++; break;
        default: h->prof.ms_staged += ms, h->prof.n_staged++; break;
        }
        h->event_pool.push_back(p.a);
        h->event_pool.push_back(p.b);
    }
    h->pending.clear();
    *out = h->prof;
    if (reset) h->prof = wrp_profile{};
    return WRP_OK;
}

// ---- HBM-resident batch -----------------------------------------------------------------
// mirror_off: float offset of dev_out inside the caller's product buffer (the mirrors take the same offset);
// mirror_off == NO_MIRRORS: this call's products are not mirrored (host destination, ring submissions)
static constexpr size_t NO_MIRRORS = ~(size_t)0;
static int process_device_impl(wrp_handle *h, const void *dev_iq, int n_sectors, float *dev_out,
                               cudaStream_t st, size_t mirror_off = NO_MIRRORS)
{
    const wrp_config &c = h->cfg;
    const int M = c.n_rows_M, N = c.n_cols_N, C = c.n_channels;
    const size_t in_bytes = input_bytes_per_sector(c);
    const size_t out_floats = (size_t)M; // 2 * M/2
    for (int s0 = 0; s0 < n_sectors; s0 += h->chunk) {
        const int S = n_sectors - s0 < h->chunk ? n_sectors - s0 : h->chunk;
        const uint8_t *in = (const uint8_t *)dev_iq + (size_t)s0 * in_bytes;
        float *out = dev_out + (size_t)s0 * out_floats;
        const int n_mirrors = mirror_off == NO_MIRRORS ? 0 : h->n_mirrors;
        bool mirrored_by_kernel = false;
        if (c.mode == WRP_MODE_STAGED) {
            unsigned long long n = 0;
            {
                ProfScope ps(h, st, 3);
                CK(h, wrp::run_staged(h, in, S, out, st, &n));
            }
            h->launches += n;
        } else {
            const void *chain_in = in;
            if (h->decode_prepass) {
                ProfScope ps(h, st, 0);
                CK(h, wrp::launch_decode_wire(in, h->decoded, M, N, C, S, st));
                h->launches++;
                chain_in = h->decoded;
            }
            if (h->chain == wrp_handle::CHAIN_STREAM) {
                ProfScope ps(h, st, 4);
                wrp::StreamParams p{};
                p.wrc_t = h->fused.wrc_t;
                p.tw_a = h->fused.tw_a;
                p.wd = h->fused.wd;
                p.wr4 = h->fused.wr4;
                p.tw4 = h->fused.tw4;
                p.tile_tw = h->fused.tile_tw;
                p.in = chain_in;
                p.out = out;
                for (int m = 0; m < n_mirrors; m++) p.mirror[m] = h->mirrors[m] + mirror_off + (size_t)s0 * out_floats;
                p.n_mirrors = n_mirrors;
                mirrored_by_kernel = true;
                p.power = h->power;
                p.x2_tap = h->x2_tap ? h->x2_tap + (size_t)s0 * C * (M / 2) * N : nullptr;
                p.scratch = h->stream_scratch;
                p.plane_cnt = h->stream_cnt;
                p.sector_cnt = h->stream_cnt + (size_t)S * C;
                p.S = S;
                p.C = C;
                p.N = N;
                p.range_res = c.range_res_m;
                p.calib = c.calib;
                p.taps_sum = h->host.taps_sum;
                memcpy(p.wcol, h->wcol, sizeof p.wcol);
                const int wire_direct = c.input_fmt == WRP_FMT_WIRE_I16BE && !h->decode_prepass;
                CUtensorMap tmap{};
                if (h->wire3) {
                    if (!wrp::wire3_encode_tensor_map(h->tma_encode, &tmap, chain_in, N, S))
                        return fail(h, WRP_ERR_CUDA, "wrp_process_device: cuTensorMapEncodeTiled rejected the wire batch (is the device buffer 16-byte aligned?)");
                    CK(h, wrp::launch_wire3(p, h->stream_max_grid, h->sm_count, tmap, st));
                    h->launches++;
                    h->prof.sectors += h->profiling ? S : 0;
                    continue;
                }
                if (!wire_direct && !wrp::stream_encode_tensor_map(h->tma_encode, &tmap, chain_in, M, N, (long long)S * C, h->l2_promotion))
                    return fail(h, WRP_ERR_CUDA, "wrp_process_device: cuTensorMapEncodeTiled rejected the batch (is the device buffer 16-byte aligned?)");
                CK(h, wrp::launch_stream(p, M, wire_direct, h->stream_max_grid, h->sm_count, (c.debug & 32) != 0, tmap, st));
                h->launches++;
            } else if (h->chain == wrp_handle::CHAIN_QUEUE) {
                ProfScope ps(h, st, 4);
                CK(h, wrp::launch_persistent((const float2 *)chain_in, out, nullptr, h->x2, h->x2_ring, h->x2_lag, h->ctrl,
                                             h->smax, h->fused, M, N, C, S, c.range_res_m, c.calib, h->host.taps_sum,
                                             h->sm_count, c.doppler_form == WRP_DOPPLER_FFT, c.evict_first, c.debug, st));
                h->launches++;
                if (c.debug & 16) {
                    int na = 0, nb = 0;
                    cudaStreamSynchronize(st);
                    wrp::persistent_debug_counters(h->ctrl, &na, &nb);
                    fprintf(stderr, "[wrp debug] %d sectors: dependency unmet at probe: %d range tiles, %d Doppler blocks\n", S, na, nb);
                }
            } else {
                {
                    ProfScope ps(h, st, 1);
                    CK(h, wrp::launch_range_fft((const float2 *)chain_in, h->x2, h->fused, M, N, C, S, st));
                    h->launches++;
                }
                {
                    ProfScope ps(h, st, 2);
                    CK(h, wrp::launch_doppler(h->x2, out, h->power, h->fused, M, N, C, S, c.range_res_m, c.calib,
                                              h->host.taps_sum, st));
                    h->launches++;
                }
            }
        }
        if (!mirrored_by_kernel) // the other kernels: a copy per mirror behind them on the same stream
            for (int m = 0; m < n_mirrors; m++)
                CK(h, cudaMemcpyAsync(h->mirrors[m] + mirror_off + (size_t)s0 * out_floats, out,
                                      (size_t)S * out_floats * sizeof(float), cudaMemcpyDefault, st));
        h->prof.sectors += h->profiling ? S : 0;
    }
    return WRP_OK;
}

int wrp_process_device(wrp_handle *h, const void *dev_iq, int n_sectors, float *dev_out, void *cuda_stream)
{
    if (!h) return WRP_ERR_INVALID;
    NvtxRange r("wrp_process_device");
    if (n_sectors < 0) return fail(h, WRP_ERR_INVALID, "wrp_process_device: negative n_sectors");
    if (n_sectors == 0) return WRP_OK;
    if (!dev_iq || !dev_out) return fail(h, WRP_ERR_INVALID, "wrp_process_device: NULL buffer");
    CK(h, cudaSetDevice(h->device));
    return process_device_impl(h, dev_iq, n_sectors, dev_out, (cudaStream_t)cuda_stream, 0);
}

// ---- pinned ring / host path ----------------------------------------------------------------
static int ensure_slot(wrp_handle *h, wrp::RingSlot &s)
{
    if (s.stream) return WRP_OK;
    const wrp_config &c = h->cfg;
    const size_t in_bytes = input_bytes_per_sector(c) * c.max_batch;
    const size_t out_bytes = (size_t)c.n_rows_M * sizeof(float) * c.max_batch;
    CK(h, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    CK(h, cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    CK(h, cudaMalloc(&s.dev_in, in_bytes));
    CK(h, cudaMalloc((void **)&s.dev_out, out_bytes));
    CK(h, cudaMallocHost((void **)&s.pinned_out, out_bytes));
    return WRP_OK;
}

static bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// enqueue one ring slot: H2D on the slot's copy stream, kernels on the compute stream,
// D2H back on the slot stream; slot.done fires when the products are in pinned_out.
static int enqueue_slot(wrp_handle *h, wrp::RingSlot &s, const void *host_iq, int n, float *dev_dst = nullptr,
                        size_t mirror_off = NO_MIRRORS)
{
    const wrp_config &c = h->cfg;
    const size_t in_bytes = input_bytes_per_sector(c) * (size_t)n;
    const void *src = host_iq;
    if (!host_ptr_is_pinned(host_iq)) {
        // pageable source: stage through the slot's pinned buffer (allocated on first use)
        if (!s.pinned_in) CK(h, cudaMallocHost(&s.pinned_in, input_bytes_per_sector(c) * c.max_batch));
        memcpy(s.pinned_in, host_iq, in_bytes);
        src = s.pinned_in;
    }
    CK(h, cudaMemcpyAsync(s.dev_in, src, in_bytes, cudaMemcpyHostToDevice, s.stream));
    CK(h, cudaEventRecord(s.h2d_done, s.stream));
    CK(h, cudaStreamWaitEvent(h->compute_stream, s.h2d_done, 0));
    const int rc = process_device_impl(h, s.dev_in, n, dev_dst ? dev_dst : s.dev_out, h->compute_stream,
                                       dev_dst ? mirror_off : NO_MIRRORS);
    if (rc != WRP_OK) return rc;
    CK(h, cudaEventRecord(s.done, h->compute_stream));
    if (!dev_dst) { // products back to the host through the slot's pinned buffer
        CK(h, cudaStreamWaitEvent(s.stream, s.done, 0));
        CK(h, cudaMemcpyAsync(s.pinned_out, s.dev_out, (size_t)n * c.n_rows_M * sizeof(float),
                              cudaMemcpyDeviceToHost, s.stream));
        CK(h, cudaEventRecord(s.done, s.stream));
    }
    s.n_sectors = n;
    return WRP_OK;
}

int wrp_submit(wrp_handle *h, const void *host_iq, int n_sectors, const int *sector_ids, const int *elev_ids)
{
    if (!h) return WRP_ERR_INVALID;
    NvtxRange r("wrp_submit");
    if (n_sectors < 1 || n_sectors > h->cfg.max_batch || !host_iq)
        return fail(h, WRP_ERR_INVALID, "wrp_submit: n_sectors must be in [1, max_batch] and host_iq non-NULL");
    if (h->ring_inflight == (int)h->ring.size())
        return fail(h, WRP_ERR_FULL, "wrp_submit: ring full, call wrp_collect");
    CK(h, cudaSetDevice(h->device));
    wrp::RingSlot &s = h->ring[h->ring_head];
    int rc = ensure_slot(h, s);
    if (rc != WRP_OK) return rc;
    rc = enqueue_slot(h, s, host_iq, n_sectors);
    if (rc != WRP_OK) return rc;
    s.sector_ids.assign(n_sectors, 0);
    s.elev_ids.assign(n_sectors, 0);
    for (int i = 0; i < n_sectors; i++) {
        s.sector_ids[i] = sector_ids ? sector_ids[i] : i;
        s.elev_ids[i] = elev_ids ? elev_ids[i] : 0;
    }
    h->ring_head = (h->ring_head + 1) % (int)h->ring.size();
    h->ring_inflight++;
    return WRP_OK;
}

int wrp_collect(wrp_handle *h, float *out_zdb_zdr, int *sector_ids, int *elev_ids, int capacity_sectors,
                int *n_done)
{
    if (!h || !n_done) return WRP_ERR_INVALID;
    NvtxRange r("wrp_collect");
    *n_done = 0;
    if (h->ring_inflight == 0) return WRP_OK;
    wrp::RingSlot &s = h->ring[h->ring_tail];
    if (capacity_sectors < s.n_sectors || !out_zdb_zdr)
        return fail(h, WRP_ERR_INVALID, "wrp_collect: output capacity smaller than the oldest submission");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaEventSynchronize(s.done));
    memcpy(out_zdb_zdr, s.pinned_out, (size_t)s.n_sectors * h->cfg.n_rows_M * sizeof(float));
    for (int i = 0; i < s.n_sectors; i++) {
        if (sector_ids) sector_ids[i] = s.sector_ids[i];
        if (elev_ids) elev_ids[i] = s.elev_ids[i];
    }
    *n_done = s.n_sectors;
    s.n_sectors = 0;
    h->ring_tail = (h->ring_tail + 1) % (int)h->ring.size();
    h->ring_inflight--;
    return WRP_OK;
}

// host_out != NULL: products to host memory; dev_out != NULL: products stay on the device
static int process_host_impl(wrp_handle *h, const void *host_iq, int n_sectors, float *host_out, float *dev_out)
{
    CK(h, cudaSetDevice(h->device));
    const wrp_config &c = h->cfg;
    const size_t in_bytes = input_bytes_per_sector(c);
    const size_t out_floats = (size_t)c.n_rows_M;
    const int depth = (int)h->ring.size();
    struct Piece {
        int first, n;
    };
    std::vector<Piece> inflight(depth, Piece{0, 0});
    int piece = 0;
    auto retire = [&](int slot) -> int {
        wrp::RingSlot &s = h->ring[slot];
        if (inflight[slot].n == 0) return WRP_OK;
        CK(h, cudaEventSynchronize(s.done));
        if (host_out)
            memcpy(host_out + (size_t)inflight[slot].first * out_floats, s.pinned_out,
                   (size_t)inflight[slot].n * out_floats * sizeof(float));
        inflight[slot].n = 0;
        s.n_sectors = 0;
        return WRP_OK;
    };
    for (int s0 = 0; s0 < n_sectors; s0 += c.max_batch, ++piece) {
        const int n = n_sectors - s0 < c.max_batch ? n_sectors - s0 : c.max_batch;
        const int slot = piece % depth;
        int rc = retire(slot);
        if (rc != WRP_OK) return rc;
        wrp::RingSlot &s = h->ring[slot];
        rc = ensure_slot(h, s);
        if (rc != WRP_OK) return rc;
        rc = enqueue_slot(h, s, (const uint8_t *)host_iq + (size_t)s0 * in_bytes, n,
                          dev_out ? dev_out + (size_t)s0 * out_floats : nullptr, (size_t)s0 * out_floats);
        if (rc != WRP_OK) return rc;
        inflight[slot] = Piece{s0, n};
    }
    for (int k = 0; k < depth; k++) {
        const int rc = retire((piece + k) % depth);
        if (rc != WRP_OK) return rc;
    }
    return WRP_OK;
}

int wrp_process_host(wrp_handle *h, const void *host_iq, int n_sectors, float *host_out)
{
    if (!h) return WRP_ERR_INVALID;
    NvtxRange r("wrp_process_host");
    if (n_sectors < 0) return fail(h, WRP_ERR_INVALID, "wrp_process_host: negative n_sectors");
    if (n_sectors == 0) return WRP_OK;
    if (!host_iq || !host_out) return fail(h, WRP_ERR_INVALID, "wrp_process_host: NULL buffer");
    if (h->ring_inflight) return fail(h, WRP_ERR_STATE, "wrp_process_host: submissions pending, collect them first");
    return process_host_impl(h, host_iq, n_sectors, host_out, nullptr);
}

int wrp_process_host_to_device(wrp_handle *h, const void *host_iq, int n_sectors, float *dev_out)
{
    if (!h) return WRP_ERR_INVALID;
    NvtxRange r("wrp_process_host_to_device");
    if (n_sectors < 0) return fail(h, WRP_ERR_INVALID, "wrp_process_host_to_device: negative n_sectors");
    if (n_sectors == 0) return WRP_OK;
    if (!host_iq || !dev_out) return fail(h, WRP_ERR_INVALID, "wrp_process_host_to_device: NULL buffer");
    if (h->ring_inflight) return fail(h, WRP_ERR_STATE, "wrp_process_host_to_device: submissions pending, collect them first");
    return process_host_impl(h, host_iq, n_sectors, nullptr, dev_out);
}

int wrp_alloc_pinned(size_t bytes, void **out)
{
    if (!out) return WRP_ERR_INVALID;
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) {
        g_create_error = std::string("wrp_alloc_pinned: ") + cudaGetErrorString(e);
        cudaGetLastError();
        *out = nullptr;
        return e == cudaErrorMemoryAllocation ? WRP_ERR_NOMEM : WRP_ERR_CUDA;
    }
    return WRP_OK;
}

int wrp_free_pinned(void *p)
{
    if (!p) return WRP_OK;
    return cudaFreeHost(p) == cudaSuccess ? WRP_OK : WRP_ERR_CUDA;
}

// ---- stage dumps -------------------------------------------------------------------------------
int wrp_dump_stage(wrp_handle *h, int sector_in_batch, int stage, int channel, void *host_out, size_t *bytes)
{
    if (!h || !bytes) return WRP_ERR_INVALID;
    const wrp_config &c = h->cfg;
    if (c.mode != WRP_MODE_STAGED) return fail(h, WRP_ERR_STATE, "wrp_dump_stage: handle is not in WRP_MODE_STAGED");
    wrp::StagedBuffers &b = h->staged;
    if (sector_in_batch < 0 || sector_in_batch >= b.last_batch)
        return fail(h, WRP_ERR_STATE, "wrp_dump_stage: no staged run holds that sector");
    if (channel < 0 || channel >= c.n_channels) return fail(h, WRP_ERR_INVALID, "wrp_dump_stage: bad channel");
    const int M = c.n_rows_M, N = c.n_cols_N, C = c.n_channels;
    const size_t mn = (size_t)M * N, hmn = (size_t)(M / 2) * N;
    const size_t plane = (size_t)sector_in_batch * C + channel;
    const void *src = nullptr;
    size_t n = 0;
    switch (stage) {
    case WRP_STAGE_00_IQ: src = b.s00 + plane * mn, n = mn * 8; break;
    case WRP_STAGE_01_HAMM: src = b.s01 + plane * mn, n = mn * 8; break;
    case WRP_STAGE_02_FFT1: src = b.s02 + plane * mn, n = mn * 8; break;
    case WRP_STAGE_03_FFT2: src = b.s03 + plane * mn, n = mn * 8; break;
    case WRP_STAGE_04_ABS: src = b.s04 + plane * hmn, n = hmn * 4; break;
    case WRP_STAGE_05_FFT3: src = b.s05 + plane * hmn, n = hmn * 8; break;
    case WRP_STAGE_06_MULT: src = b.s06 + plane * hmn, n = hmn * 8; break;
    case WRP_STAGE_07_CONV: src = b.s07 + plane * hmn, n = hmn * 8; break;
    case WRP_STAGE_08_POW: src = b.s08 + plane * hmn, n = hmn * 4; break;
    case WRP_STAGE_POWER: src = b.power + plane * (M / 2), n = (size_t)(M / 2) * 4; break;
    case WRP_STAGE_09_ZDB:
    case WRP_STAGE_10_ZDR: n = (size_t)(M / 2) * 4; break;
    default: return fail(h, WRP_ERR_INVALID, "wrp_dump_stage: unknown stage");
    }
    *bytes = n;
    if (!host_out) return WRP_OK;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    if (stage == WRP_STAGE_09_ZDB || stage == WRP_STAGE_10_ZDR) {
        // de-interleave the result slot [gate][2] of that sector
        std::vector<float> r((size_t)M);
        CK(h, cudaMemcpy(r.data(), b.result + (size_t)sector_in_batch * M, r.size() * 4, cudaMemcpyDeviceToHost));
        float *o = (float *)host_out;
        for (int g = 0; g < M / 2; g++) o[g] = r[2 * g + (stage == WRP_STAGE_10_ZDR ? 1 : 0)];
        return WRP_OK;
    }
    CK(h, cudaMemcpy(host_out, src, n, cudaMemcpyDeviceToHost));
    return WRP_OK;
}

int wrp_pack_products(const float *zdb_zdr, int gates, int sector, int elev, int with_elev, uint8_t *zdb_packet,
                      uint8_t *zdr_packet)
{
    if (!zdb_zdr || gates < 0 || (!zdb_packet && !zdr_packet)) return -WRP_ERR_INVALID;
    const int hdr = with_elev ? 4 : 2;
    uint8_t *pk[2] = {zdb_packet, zdr_packet};
    for (int which = 0; which < 2; which++) {
        uint8_t *p = pk[which];
        if (!p) continue;
        p[0] = (uint8_t)((sector >> 8) & 0xff);
        p[1] = (uint8_t)(sector & 0xff);
        if (with_elev) {
            p[2] = (uint8_t)((elev >> 8) & 0xff);
            p[3] = (uint8_t)(elev & 0xff);
        }
        for (int g = 0; g < gates; g++) {
            uint32_t u;
            memcpy(&u, &zdb_zdr[2 * g + which], 4);
            uint8_t *q = p + hdr + 4 * g;
            q[0] = (uint8_t)(u >> 24);
            q[1] = (uint8_t)(u >> 16);
            q[2] = (uint8_t)(u >> 8);
            q[3] = (uint8_t)u;
        }
    }
    return hdr + 4 * gates;
}

} // extern "C"
