// wrp_volume.cpp — a volume scan over several devices of one box behind the C ABI (include/wrp.h,
// wrp_volume_*): the reference's (elevation, sector) walk (advance(), rpv2.cu:572-579; result slot
// sitdim(0, 0, sector, elevation), rpv2.cu:607, 736) cut into one contiguous unit block per device.
//
// One host thread per shard, each with its own libwrp handle (pinned ring, copy streams, compute
// stream) on its device: H2D of the shard's sectors overlaps compute.  The only exchange is the
// gather of the finished products, and it is fused into the chain kernel: the volume buffer lives on
// devices[0], every other device has peer access to it, and each shard's handle carries its slice
// of that buffer as a product mirror (wrp_set_product_mirrors) — the kernel's epilogue stores the
// products over NVLink itself.  Shard 0 writes its slice directly.  One D2H copy returns the volume.
// A device without peer access to devices[0] falls back to cudaMemcpyPeerAsync of its slice.
// No collective library is involved and nothing here waits on another device's kernel.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wrp.h"

struct wrp_volume {
    wrp_config cfg{};
    int n_sectors = 0, n_elevations = 0;
    std::vector<int> devices;
    std::vector<wrp_handle *> handles;
    std::vector<float *> dev_out;    // per shard: [n_units of the shard][M/2][2] on its device
    std::vector<cudaStream_t> stream; // per shard: the gather copy (devices without peer access only)
    std::vector<char> peer;          // per shard: can store into dev_volume directly
    float *dev_volume = nullptr;     // on devices[0]: [U][M/2][2]
    std::string err;
};

static std::string g_volume_error;

static void shard_bounds(int units, int shard, int shards, int *lo, int *hi)
{
    // ceil(U g / G): the same rule as volume.py::shard_bounds
    *lo = (int)(((long long)units * shard + shards - 1) / shards);
    *hi = (int)(((long long)units * (shard + 1) + shards - 1) / shards);
}

static size_t sector_in_bytes(const wrp_config &c)
{
    const size_t mn = (size_t)c.n_rows_M * c.n_cols_N;
    return c.input_fmt == WRP_FMT_WIRE_I16BE ? mn * 12 : mn * 8 * c.n_channels;
}

extern "C" {

const char *wrp_volume_last_error(const wrp_volume *v) { return v ? v->err.c_str() : g_volume_error.c_str(); }

void wrp_volume_destroy(wrp_volume *v)
{
    if (!v) return;
    for (size_t g = 0; g < v->handles.size(); g++) {
        cudaSetDevice(v->devices[g]);
        if (v->handles[g]) wrp_destroy(v->handles[g]);
        if (g < v->dev_out.size() && v->dev_out[g]) cudaFree(v->dev_out[g]);
        if (g < v->stream.size() && v->stream[g]) cudaStreamDestroy(v->stream[g]);
    }
    if (v->dev_volume) {
        cudaSetDevice(v->devices[0]);
        cudaFree(v->dev_volume);
    }
    delete v;
}

int wrp_volume_create(const wrp_config *cfg, const int *devices, int n_devices, int n_sectors, int n_elevations,
                      wrp_volume **out)
{
    if (!cfg || !devices || !out || n_devices < 1 || n_devices > 64 || n_sectors < 1 || n_elevations < 1) {
        g_volume_error = "wrp_volume_create: bad argument";
        return WRP_ERR_INVALID;
    }
    *out = nullptr;
    wrp_volume *v = new wrp_volume();
    v->cfg = *cfg;
    v->n_sectors = n_sectors;
    v->n_elevations = n_elevations;
    v->devices.assign(devices, devices + n_devices);
    v->handles.assign(n_devices, nullptr);
    v->dev_out.assign(n_devices, nullptr);
    v->stream.assign(n_devices, nullptr);
    v->peer.assign(n_devices, 0);
    const int U = n_sectors * n_elevations;
    const size_t slot = (size_t)cfg->n_rows_M * sizeof(float); // 2 * M/2 floats per unit
    auto bail = [&](int rc, const std::string &msg) {
        g_volume_error = msg;
        wrp_volume_destroy(v);
        return rc;
    };
    for (int g = 0; g < n_devices; g++) {
        int lo, hi;
        shard_bounds(U, g, n_devices, &lo, &hi);
        const int rc = wrp_create(cfg, devices[g], &v->handles[g]);
        if (rc != WRP_OK) return bail(rc, std::string("wrp_volume_create: ") + wrp_last_error(nullptr));
        if (cudaSetDevice(devices[g]) != cudaSuccess ||
            cudaMalloc((void **)&v->dev_out[g], slot * (size_t)(hi - lo > 0 ? hi - lo : 1)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&v->stream[g], cudaStreamNonBlocking) != cudaSuccess)
            return bail(WRP_ERR_CUDA, std::string("wrp_volume_create: ") + cudaGetErrorString(cudaGetLastError()));
        v->peer[g] = devices[g] == devices[0];
        if (!v->peer[g]) { // NVLink path for the fused gather where the devices are peers
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[g], devices[0]) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) v->peer[g] = 1;
                cudaGetLastError(); // "already enabled" is fine
            }
        }
    }
    if (cudaSetDevice(devices[0]) != cudaSuccess || cudaMalloc((void **)&v->dev_volume, slot * (size_t)U) != cudaSuccess)
        return bail(WRP_ERR_NOMEM, "wrp_volume_create: cannot allocate the volume buffer on devices[0]");
    *out = v;
    return WRP_OK;
}

int wrp_volume_shard(const wrp_volume *v, int shard, int *first_unit, int *n_units)
{
    if (!v || shard < 0 || shard >= (int)v->devices.size() || !first_unit || !n_units) return WRP_ERR_INVALID;
    int lo, hi;
    shard_bounds(v->n_sectors * v->n_elevations, shard, (int)v->devices.size(), &lo, &hi);
    *first_unit = lo;
    *n_units = hi - lo;
    return WRP_OK;
}

int wrp_volume_process(wrp_volume *v, const void *host_iq, float *host_volume)
{
    if (!v) return WRP_ERR_INVALID;
    if (!host_iq || !host_volume) {
        v->err = "wrp_volume_process: NULL buffer";
        return WRP_ERR_INVALID;
    }
    const int G = (int)v->devices.size(), U = v->n_sectors * v->n_elevations;
    const size_t in_bytes = sector_in_bytes(v->cfg);
    const size_t slot = (size_t)v->cfg.n_rows_M * sizeof(float);
    std::vector<int> rc(G, WRP_OK);
    std::vector<std::string> msg(G);
    auto work = [&](int g) {
        int lo, hi;
        shard_bounds(U, g, G, &lo, &hi);
        if (hi <= lo) return;
        if (cudaSetDevice(v->devices[g]) != cudaSuccess) {
            rc[g] = WRP_ERR_CUDA, msg[g] = "cudaSetDevice failed";
            return;
        }
        float *slice = (float *)((uint8_t *)v->dev_volume + (size_t)lo * slot);
        const uint8_t *src = (const uint8_t *)host_iq + (size_t)lo * in_bytes;
        if (g == 0 || v->devices[g] == v->devices[0]) { // the volume's own device: straight into the slice
            rc[g] = wrp_process_host_to_device(v->handles[g], src, hi - lo, slice);
            if (rc[g] != WRP_OK) msg[g] = wrp_last_error(v->handles[g]);
            return;
        }
        if (v->peer[g]) { // fused gather: the kernel's epilogue stores the slice over NVLink
            rc[g] = wrp_set_product_mirrors(v->handles[g], &slice, 1);
            if (rc[g] == WRP_OK) rc[g] = wrp_process_host_to_device(v->handles[g], src, hi - lo, v->dev_out[g]);
            if (rc[g] != WRP_OK) msg[g] = wrp_last_error(v->handles[g]);
            return;
        }
        rc[g] = wrp_process_host_to_device(v->handles[g], src, hi - lo, v->dev_out[g]);
        if (rc[g] != WRP_OK) {
            msg[g] = wrp_last_error(v->handles[g]);
            return;
        }
        // no peer access: this shard's slice of the product volume -> devices[0] by DMA
        cudaError_t e = cudaMemcpyPeerAsync(slice, v->devices[0], v->dev_out[g], v->devices[g], (size_t)(hi - lo) * slot,
                                            v->stream[g]);
        if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream[g]);
        if (e != cudaSuccess) rc[g] = WRP_ERR_CUDA, msg[g] = std::string("gather: ") + cudaGetErrorString(e);
    };
    std::vector<std::thread> threads;
    for (int g = 1; g < G; g++) threads.emplace_back(work, g);
    work(0);
    for (auto &t : threads) t.join();
    for (int g = 0; g < G; g++)
        if (rc[g] != WRP_OK) {
            v->err = "wrp_volume_process: shard " + std::to_string(g) + ": " + msg[g];
            return rc[g];
        }
    if (cudaSetDevice(v->devices[0]) != cudaSuccess ||
        cudaMemcpy(host_volume, v->dev_volume, (size_t)U * slot, cudaMemcpyDeviceToHost) != cudaSuccess) {
        v->err = std::string("wrp_volume_process: volume D2H: ") + cudaGetErrorString(cudaGetLastError());
        return WRP_ERR_CUDA;
    }
    return WRP_OK;
}

} // extern "C"
