// wrp_persistent.cu — the fused chain as ONE persistent kernel per batch (sm_100a).
//
// Work items, handed out in queue order by an atomic counter:
//   A(s, tile)  range tile: 8 adjacent columns x 1024 rows of one (sector, channel) plane.
//               TMA (cp.async.bulk.tensor, one 64 KiB 4-D box) -> window on load -> radix-32 x
//               radix-32 column FFT with an in-place shared-memory exchange -> rows k < M/2 to
//               the x2 ring (stages 01-02; replaces __apply_hamming + the strided cuFFT plan,
//               rpv2.cu:86-91, 318-333, 418-428).
//   B(s, block) Doppler block: 16 rows (8 gates x hh,vv or 16 gates of vh) of the x2 ring.
//               1-D bulk copies (cp.async.bulk, 2 x 32 KiB) -> mean removal, inverse DFT, shift,
//               clip, |.|^2, moving-average power, ZdB/ZDR (stages 03-10; replaces rpv2.cu:93-213,
//               434-566).  Rows are warp-private: no CTA barrier inside the transform.
// The queue interleaves A(t) with B(t - lag), so the x2 hand-off lives in a small ring
// (ring x 6 MiB) that never leaves L2, there is no kernel boundary between the phases, and
// DRAM-heavy A items overlap compute-heavy B items on every SM.  Each CTA prefetches its next
// item's bytes into the same 64 KiB buffer as soon as the current item's last shared-memory
// read has been issued, so loads overlap the second FFT pass and the epilogue.
// Cross-CTA ordering: per-sector completion counters (release: __threadfence + atomicAdd,
// acquire: ld.acquire.gpu spin by the one thread that issues the dependent bulk copy).  An item
// only ever waits for items earlier in the queue, which are held by running CTAs: no deadlock.
#include <cstdlib>

#include "wrp_fft.cuh"
#include "wrp_internal.h"

namespace wrp {

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 16-byte async copy global -> shared (L1 bypass).  An L2 evict-first cache hint was tried here
// (cp.async ... .L2::cache_hint): ptxas 12.9 allocated an odd uniform register for the LDGSTS
// descriptor at one call site and the warp trapped with "illegal instruction", so the L2 priority
// is left at the default (st/ld eviction-priority qualifiers need 256-bit vectors on sm_100).
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// arrive on `bar` once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// release-add: orders this thread's prior writes and, through the preceding CTA barrier, the
// whole CTA's (cumulativity); no L1 invalidation, unlike __threadfence()
__device__ __forceinline__ void red_release_add(int *p)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}
// probe without the L1 invalidate an acquire load carries (thread 0's warp would sit on it): the
// counter is bumped by a release (data is in L2 before the count moves) and everything read
// under it is fetched by bulk async copies straight from L2, issued after the value is seen
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until(const int *p, int target)
{
    while (ld_acquire(p) < target) __nanosleep(100);
}

// ---- parameters ----------------------------------------------------------------------------
struct PersistParams {
    const float *wrc_t;
    const float *wd;
    const float2 *tw_a;
    const float2 *tw_b;
    const float2 *iq; // input [S][C][M][N]
    float2 *x2;   // ring [ring][C][M/2][N]
    float *out;   // [S][M/2][2]
    float *power; // optional [S][C][M/2]
    int *ctrl;    // [0] work counter; a_done at CTRL_A; b_done at CTRL_A + smax
    int S, C, N, half_m;
    int ring, lag;
    int tiles_a, blocks_b, pair_blocks;
    int tile_bytes;
    int bulk_piece;
    int dephase_ns;
    int b_async; // Doppler blocks loaded with cp.async instead of bulk copies
    int debug; // WRP_DEBUG bisect switches (development only)
    int n1, n2, n3, b3_first; // queue regions (see decode_item)
    int total_items;
    int smax;
    float range_res, calib, taps_sum;
};
constexpr int CTRL_A = 32;
// shared-memory copies of the small tables (no L1 dependence: the acquire loads of the
// dependency counters invalidate L1)
// rows are padded by 16 B so that the 128-bit reads of lanes holding different rows hit
// different banks
constexpr int WRC_ROW = 32 * 8 + 16;   // wr(i)*c transposed [32][32], each value stored twice (w, w) for FMUL2
constexpr int TWA_ROW = 32 * 8 + 16;   // range inter-pass twiddles [32][32] float2
constexpr int TAB_WRC = 32 * WRC_ROW;
constexpr int TAB_TWA = 32 * TWA_ROW;
// T = columns per range tile = warps per CTA; the tile buffer holds T*8 KiB
template <int R1B, int T> struct Tables {
    static constexpr int TILE_BYTES = T * 8192;
    static constexpr int N = 32 * R1B;
    static constexpr int TWB_ROW = R1B * 8 + 16; // Doppler inter-pass twiddles [32][R1B] float2
    static constexpr int WD = N * 4;             // Doppler window
    static constexpr int TWB = 32 * TWB_ROW;
    static constexpr int OFF_WRC = TILE_BYTES;
    static constexpr int OFF_TWA = OFF_WRC + TAB_WRC;
    static constexpr int OFF_TWB = OFF_TWA + TAB_TWA;
    static constexpr int OFF_WD = OFF_TWB + TWB;
    static constexpr int SMEM = OFF_WD + WD;
};

struct Item {
    int kind; // 0 = range tile, 1 = Doppler block
    int sector;
    int sub;
};

// queue: n1 steps of [A]; n2 steps of [A, B]; n3 steps of [B]
__device__ __forceinline__ Item decode_item(int idx, const PersistParams &p)
{
    Item it;
    const int TA = p.tiles_a, TB = p.blocks_b;
    if (idx < p.n1 * TA) {
        it.kind = 0;
        it.sector = idx / TA;
        it.sub = idx - it.sector * TA;
        return it;
    }
    idx -= p.n1 * TA;
    const int per = TA + TB;
    if (idx < p.n2 * per) {
        const int t = idx / per, r = idx - t * per;
        if (r < TA) {
            it.kind = 0;
            it.sector = p.n1 + t;
            it.sub = r;
        } else {
            it.kind = 1;
            it.sector = t;
            it.sub = r - TA;
        }
        return it;
    }
    idx -= p.n2 * per;
    it.kind = 1;
    const int t = idx / TB;
    it.sector = p.b3_first + t;
    it.sub = idx - t * TB;
    return it;
}

// dependency of an item: the counter it must see reach `target` before its copy may start
//   range tile of sector s >= ring: WAR on ring slot — the Doppler blocks of sector s - ring must
//                                   have pulled their rows;
//   Doppler block of sector s:      every range tile of sector s has been written.
__device__ __forceinline__ const int *item_dep(const Item &it, const PersistParams &p, int &target)
{
    if (it.kind == 0) {
        target = p.blocks_b;
        return it.sector >= p.ring ? p.ctrl + CTRL_A + p.smax + (it.sector - p.ring) : nullptr;
    }
    target = p.tiles_a;
    return p.ctrl + CTRL_A + it.sector;
}
__device__ __forceinline__ bool dep_ready(const Item &it, const PersistParams &p)
{
    int target;
    const int *dep = item_dep(it, p, target);
    return dep == nullptr || ld_relaxed(dep) >= target;
}
__device__ __forceinline__ void dep_wait(const Item &it, const PersistParams &p)
{
    int target;
    const int *dep = item_dep(it, p, target);
    if (dep) spin_until(dep, target);
}

// Doppler block (thread 0): two 32 KiB (or one 64 KiB) contiguous bulk copies from the x2 ring
__device__ __forceinline__ void issue_load_b(const Item &it, const PersistParams &p, uint8_t *buf, uint64_t *bar)
{
    const int TILE_BYTES = p.tile_bytes;
    fence_proxy_async();
    mbar_expect_tx(bar, TILE_BYTES);
    const int slot = it.sector % p.ring;
    const size_t row_bytes = (size_t)p.N * sizeof(float2);
    const int rows = TILE_BYTES / (int)row_bytes;
    if (it.sub < p.pair_blocks) {
        const int g0 = it.sub * (rows / 2);
        const uint8_t *hh = (const uint8_t *)p.x2 + (((size_t)slot * p.C + 0) * p.half_m + g0) * row_bytes;
        const uint8_t *vv = (const uint8_t *)p.x2 + (((size_t)slot * p.C + 1) * p.half_m + g0) * row_bytes;
        const int piece = p.bulk_piece; // bytes per bulk request (several requests pipeline in the TMA unit)
        for (int o = 0; o < TILE_BYTES / 2; o += piece) {
            bulk_load(buf + o, hh + o, piece, bar);
            bulk_load(buf + TILE_BYTES / 2 + o, vv + o, piece, bar);
        }
    } else {
        const int g0 = (it.sub - p.pair_blocks) * rows;
        const int ch = p.C == 1 ? 0 : 2;
        const uint8_t *src = (const uint8_t *)p.x2 + (((size_t)slot * p.C + ch) * p.half_m + g0) * row_bytes;
        const int piece = p.bulk_piece;
        for (int o = 0; o < TILE_BYTES; o += piece) bulk_load(buf + o, src + o, piece, bar);
    }
}

// Doppler block through cp.async (all threads): same bytes as issue_load_b, more requests in flight
template <int T>
__device__ __forceinline__ void issue_load_b_async(const Item &it, const PersistParams &p, uint8_t *buf,
                                                   uint64_t *bar, int tid)
{
    constexpr int TILE_BYTES = T * 8192, THREADS = 32 * T;
    const int slot = it.sector % p.ring;
    const size_t row_bytes = (size_t)p.N * sizeof(float2);
    const int rows = TILE_BYTES / (int)row_bytes;
    const uint8_t *s0, *s1;
    if (it.sub < p.pair_blocks) {
        const int g0 = it.sub * (rows / 2);
        s0 = (const uint8_t *)p.x2 + (((size_t)slot * p.C + 0) * p.half_m + g0) * row_bytes;
        s1 = (const uint8_t *)p.x2 + (((size_t)slot * p.C + 1) * p.half_m + g0) * row_bytes;
    } else {
        const int g0 = (it.sub - p.pair_blocks) * rows;
        const int ch = p.C == 1 ? 0 : 2;
        s0 = (const uint8_t *)p.x2 + (((size_t)slot * p.C + ch) * p.half_m + g0) * row_bytes;
        s1 = s0 + TILE_BYTES / 2;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int o = (tid + k * THREADS) * 16;
        cp_async16(buf + o, s0 + o);
        cp_async16(buf + TILE_BYTES / 2 + o, s1 + o);
    }
    cp_async_arrive(bar);
}

// Range tile (all threads): 1024 rows x T*8 B, 16 B per cp.async (a warp covers 512 contiguous-row bytes).
// A 4-D TMA box was tried first: one 64 B row per request throttles the TMA unit (~5 us per tile).
template <int N, int T>
__device__ __forceinline__ void issue_load_a(const Item &it, const PersistParams &p, uint8_t *buf, uint64_t *bar,
                                             int tid)
{
    constexpr int tiles_per_plane = N / T;
    constexpr int CPR = T / 2; // 16-byte chunks per tile row
    const int ch = it.sub / tiles_per_plane, tile = it.sub - ch * tiles_per_plane;
    const uint8_t *src = (const uint8_t *)p.iq + ((size_t)(it.sector * p.C + ch) * 1024 + (tid / CPR)) * (N * 8) +
                         tile * (T * 8) + (tid % CPR) * 16;
    uint8_t *dst = buf + tid * 16;
#pragma unroll
    for (int k = 0; k < 16; ++k) cp_async16(dst + k * (64 * T * 8), src + (size_t)k * 64 * (N * 8));
    cp_async_arrive(bar);
}

// ---- the kernel ------------------------------------------------------------------------------
template <int R1B, int T> // Doppler length N = 32 * R1B; T columns per range tile = warps per CTA
__global__ void __launch_bounds__(32 * T, 16 / T)
    chain_persistent_kernel(const PersistParams p)
{
    constexpr int R = 32; // range FFT 32 x 32 (M = 1024)
    constexpr int N = 32 * R1B;
    constexpr int THREADS = 32 * T;
    using Tab = Tables<R1B, T>;
    constexpr int TILE_BYTES = Tab::TILE_BYTES;
    constexpr int PITCH = T * 8;                   // bytes per tile row
    constexpr int SW = 128 / PITCH - 1;            // row-swizzle mask of the in-place exchange
    constexpr int ROWS_B = TILE_BYTES / (N * 8);   // Doppler rows per block
    constexpr int RPW = ROWS_B / T;                // rows per warp
    static_assert(RPW == 1 || RPW == 2, "Doppler rows per warp");
    extern __shared__ __align__(1024) uint8_t tile[];
    __shared__ __align__(8) uint64_t mbar_a, mbar_b; // range tiles (256 cp.async arrivals) / Doppler blocks (tx bytes)
    __shared__ int4 s_item[2]; // next item decoded by thread 0: (kind, sector, sub, ready); kind < 0: queue empty
    __shared__ float p_row[ROWS_B];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        // copy the tables into (row-padded) shared memory, 16 B per step
        auto copy_rows = [&](int off, const void *src, int rows, int row_bytes, int row_pitch) {
            const int per_row = row_bytes / 16;
            for (int i = tid; i < rows * per_row; i += THREADS) {
                const int r = i / per_row, q = i - r * per_row;
                *reinterpret_cast<float4 *>(tile + off + r * row_pitch + q * 16) =
                    __ldg(reinterpret_cast<const float4 *>(src) + i);
            }
        };
        for (int i = tid; i < 32 * 32; i += THREADS) { // window values duplicated into (w, w) pairs
            const float w = __ldg(p.wrc_t + i);
            *reinterpret_cast<float2 *>(tile + Tab::OFF_WRC + (i >> 5) * WRC_ROW + (i & 31) * 8) = make_float2(w, w);
        }
        copy_rows(Tab::OFF_TWA, p.tw_a, 32, 32 * 8, TWA_ROW);
        copy_rows(Tab::OFF_TWB, p.tw_b, 32, R1B * 8, Tab::TWB_ROW);
        copy_rows(Tab::OFF_WD, p.wd, 1, Tab::WD, Tab::WD);
    }
    if (p.debug & 1) return;
    // experiment: de-phase the two CTAs of an SM (they start together and items are equal-length)
    if (p.dephase_ns > 0 && blockIdx.x >= gridDim.x / 2)
        for (int t = 0; t < p.dephase_ns; t += 500) __nanosleep(500);
    int pending = -1; // sector whose range tile this CTA finished but has not signalled yet
    if (tid == 0) {
        mbar_init(&mbar_a, THREADS);
        mbar_init(&mbar_b, 1);
        const int first = atomicAdd(p.ctrl, 1);
        Item f{-1, 0, 0};
        if (first < p.total_items) f = decode_item(first, p);
        s_item[0] = make_int4(f.kind, f.sector, f.sub, 1);
    }
    __syncthreads();
    Item it{s_item[0].x, s_item[0].y, s_item[0].z};
    if (p.debug & 4) return;
    if (it.kind >= 0) {
        if (tid == 0) dep_wait(it, p);
        if (it.kind == 0) {
            __syncthreads();
            issue_load_a<N, T>(it, p, tile, &mbar_a, tid);
        } else if (p.b_async) {
            __syncthreads();
            issue_load_b_async<T>(it, p, tile, &mbar_a, tid);
        } else if (tid == 0) {
            issue_load_b(it, p, tile, &mbar_b);
        }
    }
    uint32_t phase_a = 0, phase_b = 0;
    int it_count = 0;

    while (it.kind >= 0) {
        const int nslot = (it_count + 1) & 1;
        ++it_count;
        Item nit;   // next item (CTA-uniform after the buffer-release barrier)
        bool ready; // its dependency was met when probed
        // thread 0 claims the next queue slot now — the atomic's round trip hides behind this item's
        // first pass — and decodes it for everybody
        int claimed = 0;
        int4 nx = make_int4(-1, 0, 0, 1);
        if (tid == 0) claimed = atomicAdd(p.ctrl, 1);
        if (it.kind == 0 || p.b_async) {
            mbar_wait(&mbar_a, phase_a);
            phase_a ^= 1;
        } else {
            mbar_wait(&mbar_b, phase_b);
            phase_b ^= 1;
        }

        if (it.kind == 0) {
            // ================= range tile =================
            const int c = tid % T, b = tid / T;
            constexpr int tiles_per_plane = N / T;
            const int ch = it.sub / tiles_per_plane, col = (it.sub - ch * tiles_per_plane) * T + c;
            float2 v[R];
            {
                const uint8_t *src = tile + b * PITCH + c * 8;
                static_for<R>([&](auto ai) {
                    constexpr int a = decltype(ai)::value;
                    v[brev<R>(a)] = *reinterpret_cast<const float2 *>(src + a * (R * PITCH));
                });
                // stage 01 (x *= wr(i)*c*wd(j), rpv2.cu:86-91) fused into the first butterfly stage:
                // the span-1 partners of the bit-reversed network are rows a and a + 16
                const float wdj = reinterpret_cast<const float *>(tile + Tab::OFF_WD)[col];
                const float2 wd2 = make_float2(wdj, wdj), m2 = make_float2(-2.f, -2.f);
                const float4 *w4 = reinterpret_cast<const float4 *>(tile + Tab::OFF_WRC + b * WRC_ROW);
                static_for<R / 4>([&](auto qi) { // two rows a, a+1 per 128-bit table read
                    constexpr int q = decltype(qi)::value;
                    const float4 wlo = w4[q], whi = w4[q + R / 4];
                    static_for<2>([&](auto ei) {
                        constexpr int e = decltype(ei)::value;
                        constexpr int sa = brev<R>(2 * q + e); // even slot; partner row a + R/2 sits in sa + 1
                        static_assert(brev<R>(2 * q + e + R / 2) == sa + 1, "span-1 partner");
                        const float2 wl = cmul2(e ? make_float2(wlo.z, wlo.w) : make_float2(wlo.x, wlo.y), wd2);
                        const float2 wh = cmul2(e ? make_float2(whi.z, whi.w) : make_float2(whi.x, whi.y), wd2);
                        const float2 t = cmul2(v[sa + 1], wh);
                        const float2 s2 = cfma2(v[sa], wl, t); // A*wl + B*wh
                        v[sa + 1] = cfma2(t, m2, s2);          // A*wl - B*wh
                        v[sa] = s2;
                    });
                });
            }
            fft_dit_after_stage1<R, -1>(v);
            if (tid == 0 && claimed < p.total_items) {
                const Item c = decode_item(claimed, p);
                nx = make_int4(c.kind, c.sector, c.sub, (int)dep_ready(c, p));
            }
            __syncwarp();
            {
                // Z[ka][b] goes to row 32 ka + (b ^ (ka & SW)): the warp keeps its own row footprint
                // (in place), and the 128/PITCH values of ka met by one shared-memory wavefront of
                // pass 2 fall into different PITCH-byte slices of a 128-byte bank line
                const float4 *t4 = reinterpret_cast<const float4 *>(tile + Tab::OFF_TWA + b * TWA_ROW);
                uint8_t *d_sw[SW + 1];
#pragma unroll
                for (int sx = 0; sx <= SW; ++sx) d_sw[sx] = tile + (b ^ sx) * PITCH + c * 8;
                static_for<R / 2>([&](auto qi) {
                    constexpr int q = decltype(qi)::value;
                    const float4 w = t4[q];
                    const float2 y0 = q == 0 ? v[0] : cmul(v[2 * q], make_float2(w.x, w.y));
                    const float2 y1 = cmul(v[2 * q + 1], make_float2(w.z, w.w));
                    *reinterpret_cast<float2 *>(d_sw[(2 * q) & SW] + (2 * q) * (R * PITCH)) = y0;
                    *reinterpret_cast<float2 *>(d_sw[(2 * q + 1) & SW] + (2 * q + 1) * (R * PITCH)) = y1;
                });
            }
            __syncthreads();
            if (tid == 0 && pending >= 0) red_release_add(p.ctrl + CTRL_A + pending);
            pending = -1;
            const int ka = b;
            {
                // row 32 ka + (bb ^ (ka & SW)): the swizzle bits do not overlap the rest of the address
                const uint32_t off_sw = (uint32_t)(ka * (R * PITCH) + c * 8) | (uint32_t)((ka & SW) * PITCH);
                static_for<R>([&](auto bi) {
                    constexpr int bb = decltype(bi)::value;
                    v[brev<R>(bb)] = *reinterpret_cast<const float2 *>(tile + (off_sw ^ (uint32_t)(bb * PITCH)));
                });
            }
            if (tid == 0) s_item[nslot] = nx;
            __syncthreads(); // every shared-memory read of this item is done; s_next is visible
            nit = Item{s_item[nslot].x, s_item[nslot].y, s_item[nslot].z};
            ready = s_item[nslot].w != 0;
            if (nit.kind >= 0 && ready) {
                if (nit.kind == 0)
                    issue_load_a<N, T>(nit, p, tile, &mbar_a, tid);
                else if (p.b_async)
                    issue_load_b_async<T>(nit, p, tile, &mbar_a, tid);
                else if (tid == 0)
                    issue_load_b(nit, p, tile, &mbar_b);
            }
            fft_dit<R, -1>(v);
            {
                float2 *out = p.x2 + (((size_t)(it.sector % p.ring) * p.C + ch) * p.half_m + ka) * (size_t)N + col;
                static_for<R / 2>([&](auto ki) { // rows k = ka + 32 kb < M/2
                    constexpr int kb = decltype(ki)::value;
                    out[(size_t)(R * kb) * N] = v[kb];
                });
            }
            // completion of this tile is signalled (release) by thread 0 after the next CTA
            // barrier, when these stores have long drained: no fence stall here
            pending = it.sector;
        } else {
            // ================= Doppler block =================
            if (tid == 0) atomicAdd(p.ctrl + CTRL_A + p.smax + it.sector, 1); // ring rows are in smem now
            const bool pair = it.sub < p.pair_blocks;
            // warp w owns rows w and (RPW == 2) ROWS_B/2 + w: (hh, vv) of one gate in a pair block
            {
                float2 tw[R1B];
                const float4 *t4 = reinterpret_cast<const float4 *>(tile + Tab::OFF_TWB + lane * Tab::TWB_ROW);
                static_for<R1B / 2>([&](auto qi) {
                    constexpr int q = decltype(qi)::value;
                    const float4 w = t4[q];
                    tw[2 * q] = make_float2(w.x, w.y);
                    tw[2 * q + 1] = make_float2(w.z, w.w);
                });
#pragma unroll 1
                for (int rr = 0; rr < RPW; ++rr) {
                    uint8_t *row = tile + (size_t)(warp + rr * (ROWS_B / 2)) * (N * 8);
                    float2 v[R1B];
                    static_for<R1B>([&](auto ai) {
                        constexpr int a = decltype(ai)::value;
                        v[brev<R1B>(a)] = *reinterpret_cast<const float2 *>(row + (32 * a + lane) * 8);
                    });
                    fft_dit<R1B, +1>(v);
                    // mean removal (rpv2.cu:93-130): only the a-sum (ka = 0) carries the row mean
                    float sx = v[0].x, sy = v[0].y;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sx += __shfl_xor_sync(0xffffffffu, sx, o);
                        sy += __shfl_xor_sync(0xffffffffu, sy, o);
                    }
                    v[0].x -= sx * (1.f / 32.f);
                    v[0].y -= sy * (1.f / 32.f);
                    __syncwarp();
                    // Z_l[ka] -> float2 index 32 ka + (l ^ ((ka & 7) << 1)): 16-byte chunks of group ka
                    // are XOR-swizzled so pass 2's 128-bit reads are conflict-free
                    static_for<R1B>([&](auto ki) {
                        constexpr int ka = decltype(ki)::value;
                        const float2 y = ka == 0 ? v[0] : cmul(v[ka], tw[ka]);
                        *reinterpret_cast<float2 *>(row + (32 * ka + (lane ^ ((ka & 7) << 1))) * 8) = y;
                    });
                }
            }
            if (tid == 0 && claimed < p.total_items) {
                const Item c = decode_item(claimed, p);
                nx = make_int4(c.kind, c.sector, c.sub, (int)dep_ready(c, p));
            }
            __syncwarp();
            float2 u[32];
            const int rsel = RPW == 2 ? (lane >> 4) : 0;
            const int ka = RPW == 2 ? (lane & 15) : lane;
            const int my_row = warp + rsel * (ROWS_B / 2);
            {
                const uint8_t *grp = tile + (size_t)my_row * (N * 8) + ka * 256;
                const int sw = (ka & 7) * 16;
                static_for<16>([&](auto ci) {
                    constexpr int cc = decltype(ci)::value;
                    const float4 q = *reinterpret_cast<const float4 *>(grp + ((cc * 16) ^ sw));
                    u[brev<32>(2 * cc)] = make_float2(q.x, q.y);
                    u[brev<32>(2 * cc + 1)] = make_float2(q.z, q.w);
                });
            }
            if (tid == 0) s_item[nslot] = nx;
            __syncthreads();
            if (tid == 0 && pending >= 0) red_release_add(p.ctrl + CTRL_A + pending);
            pending = -1;
            nit = Item{s_item[nslot].x, s_item[nslot].y, s_item[nslot].z};
            ready = s_item[nslot].w != 0;
            if (nit.kind >= 0 && ready) {
                if (nit.kind == 0)
                    issue_load_a<N, T>(nit, p, tile, &mbar_a, tid);
                else if (p.b_async)
                    issue_load_b_async<T>(nit, p, tile, &mbar_a, tid);
                else if (tid == 0)
                    issue_load_b(nit, p, tile, &mbar_b);
            }
            fft_dit<32, +1>(u);
            // stage 03 shift + clip (rpv2.cu:137-148): the zeroed columns N-1, N-2 are bins N/2-1 =
            // (R1B-1) + R1B*15 and N/2-2; stage 04 |.|^2 and the row sum (rpv2.cu:150-157, 171-197)
            if (ka >= R1B - 2) u[15] = make_float2(0.f, 0.f); // the two clipped bins
            float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            static_for<32>([&](auto ki) { // (sum re^2, sum im^2) with one FFMA2 per bin
                constexpr int kb = decltype(ki)::value;
                acc[kb & 3] = cfma2(u[kb], u[kb], acc[kb & 3]);
            });
            const float2 a2 = cadd(cadd(acc[0], acc[1]), cadd(acc[2], acc[3]));
            float pw = a2.x + a2.y;
#pragma unroll
            for (int o = R1B / 2; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
            pw *= p.taps_sum; // stages 05-08: row sum of the circular convolution
            // (channel, gate) of my_row
            const int g0 = pair ? it.sub * (ROWS_B / 2) : (it.sub - p.pair_blocks) * ROWS_B;
            const int chn = pair ? (my_row >= ROWS_B / 2 ? 1 : 0) : (p.C == 1 ? 0 : 2);
            const int gate = pair ? g0 + (my_row % (ROWS_B / 2)) : g0 + my_row;
            if (p.power && ka == 0) p.power[((size_t)it.sector * p.C + chn) * p.half_m + gate] = pw;
            if constexpr (RPW == 2) {
                const float other = __shfl_xor_sync(0xffffffffu, pw, 16);
                if (pair && lane == 0) {
                    const float rg = (float)gate * p.range_res;
                    const float z = rg * rg * p.calib * pw;
                    reinterpret_cast<float2 *>(p.out)[(size_t)it.sector * p.half_m + gate] =
                        make_float2(10.f * log10f(z), 10.f * (log10f(pw) - log10f(other)));
                } else if (!pair && p.C == 1 && ka == 0) {
                    const float rg = (float)gate * p.range_res;
                    const float z = rg * rg * p.calib * pw;
                    reinterpret_cast<float2 *>(p.out)[(size_t)it.sector * p.half_m + gate] =
                        make_float2(10.f * log10f(z), 0.f);
                }
            } else {
                if (ka == 0) p_row[my_row] = pw;
                __syncthreads();
                if (pair && tid < ROWS_B / 2) { // N = 1024: (hh, vv) rows sit in different warps
                    const int g = g0 + tid;
                    const float p_hh = p_row[tid], p_vv = p_row[ROWS_B / 2 + tid];
                    const float rg = (float)g * p.range_res;
                    const float z = rg * rg * p.calib * p_hh;
                    reinterpret_cast<float2 *>(p.out)[(size_t)it.sector * p.half_m + g] =
                        make_float2(10.f * log10f(z), 10.f * (log10f(p_hh) - log10f(p_vv)));
                } else if (!pair && p.C == 1 && tid < ROWS_B) {
                    const int g = g0 + tid;
                    const float rg = (float)g * p.range_res;
                    const float z = rg * rg * p.calib * p_row[tid];
                    reinterpret_cast<float2 *>(p.out)[(size_t)it.sector * p.half_m + g] =
                        make_float2(10.f * log10f(z), 0.f);
                }
                __syncthreads();
            }
        }
        // the probe found the next item's dependency unmet: this CTA's own item is signalled
        // (or about to be, by its other warps), so a blocking wait is safe now
        if ((p.debug & 16) && tid == 0 && nit.kind >= 0 && !ready) atomicAdd(p.ctrl + 1 + nit.kind, 1);
        if (nit.kind < 0 || !ready) {
            // leaving the loop, or about to block on a dependency: publish the pending tile first
            // (the item waited for may be this CTA's own).  nit and ready are CTA-uniform.
            if (pending >= 0) {
                __syncthreads();
                if (tid == 0) red_release_add(p.ctrl + CTRL_A + pending);
                pending = -1;
            }
            if (nit.kind >= 0) {
                if (tid == 0) dep_wait(nit, p);
                if (nit.kind == 0) {
                    __syncthreads();
                    issue_load_a<N, T>(nit, p, tile, &mbar_a, tid);
                } else if (p.b_async) {
                    __syncthreads();
                    issue_load_b_async<T>(nit, p, tile, &mbar_a, tid);
                } else if (tid == 0) {
                    issue_load_b(nit, p, tile, &mbar_b);
                }
            }
        }
        it = nit;
    }
}

// ---- host side ---------------------------------------------------------------------------------
bool persistent_supported(int M, int N) { return M == 1024 && (N == 512 || N == 1024); }
int persistent_ctrl_ints(int smax) { return CTRL_A + 2 * smax; }

static int tile_cols()
{
    static int t = 0;
    if (!t) {
        const char *env = getenv("WRP_TILE_COLS");
        t = env && atoi(env) == 4 ? 4 : 8; // 8 columns (64-byte row segments) measured 24 % faster than 4
    }
    return t;
}

cudaError_t persistent_setup()
{
    cudaError_t e;
#define WRP_SET(R1B, T)                                                                                      \
    e = cudaFuncSetAttribute(chain_persistent_kernel<R1B, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                             Tables<R1B, T>::SMEM);                                                          \
    if (e != cudaSuccess) return e;
    WRP_SET(16, 8)
    WRP_SET(16, 4)
    WRP_SET(32, 8)
    WRP_SET(32, 4)
#undef WRP_SET
    return cudaSuccess;
}

// One launch for the whole batch.  ctrl must hold CTRL_A + 2*smax ints.
cudaError_t launch_persistent(const float2 *iq, float *out, float *power, float2 *x2_ring, int ring, int lag,
                              int *ctrl,
                              int smax, const FusedTables &t, int M, int N, int C, int n_sectors, float range_res,
                              float calib, float taps_sum, int sm_count, size_t l2_window_bytes, cudaStream_t st)
{
    if (n_sectors == 0) return cudaSuccess;
    if (!persistent_supported(M, N) || n_sectors > smax) return cudaErrorInvalidValue;
    PersistParams p{};
    p.wrc_t = t.wrc_t;
    p.wd = t.wd;
    p.tw_a = t.tw_a;
    p.tw_b = t.tw_b;
    p.iq = iq;
    p.x2 = x2_ring;
    p.out = out;
    p.power = power;
    p.ctrl = ctrl;
    p.S = n_sectors;
    p.C = C;
    p.N = N;
    p.half_m = M / 2;
    p.ring = ring;
    p.lag = lag;
    const int T = tile_cols();
    p.tile_bytes = T * 8192;
    const int rows_b = p.tile_bytes / (N * 8);
    p.tiles_a = (N / T) * C;
    p.pair_blocks = C >= 2 ? (M / 2) / (rows_b / 2) : 0;
    p.blocks_b = p.pair_blocks + ((C & 1) ? (M / 2) / rows_b : 0);
    const int L = p.lag, S = n_sectors;
    p.n1 = S < L ? S : L;
    p.n2 = S > L ? S - L : 0;
    p.n3 = S < L ? S : L;
    p.b3_first = S > L ? S - L : 0;
    p.total_items = S * (p.tiles_a + p.blocks_b);
    p.smax = smax;
    p.bulk_piece = getenv("WRP_BULK_PIECE") ? atoi(getenv("WRP_BULK_PIECE")) : p.tile_bytes / 2;
    if (p.bulk_piece < 512 || (p.tile_bytes / 2) % p.bulk_piece) p.bulk_piece = p.tile_bytes / 2;
    p.b_async = getenv("WRP_B_ASYNC") ? atoi(getenv("WRP_B_ASYNC")) : 1; // +1 % over bulk copies
    p.dephase_ns = getenv("WRP_DEPHASE_NS") ? atoi(getenv("WRP_DEPHASE_NS")) : 0;
    p.debug = getenv("WRP_DEBUG") ? atoi(getenv("WRP_DEBUG")) : 0;
    p.range_res = range_res;
    p.calib = calib;
    p.taps_sum = taps_sum;

    cudaError_t e = cudaMemsetAsync(ctrl, 0, sizeof(int) * (CTRL_A + 2 * (size_t)smax), st);
    if (e != cudaSuccess) return e;
    int grid = (16 / T) * sm_count;
    if (grid > p.total_items) grid = p.total_items;

    // Optional (WRP_L2_PERSIST=1, needs a persisting-L2 carve-out set by the caller of wrp_create):
    // pin the x2 ring in L2 for this launch through an access-policy window, so the streamed
    // input cannot push the hand-off rows out to DRAM.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(32 * T);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int n_attr = 0;
    if (l2_window_bytes > 0) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = (void *)x2_ring;
        attr[0].val.accessPolicyWindow.num_bytes = l2_window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        n_attr = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    if (N == 512 && T == 8) {
        cfg.dynamicSmemBytes = Tables<16, 8>::SMEM;
        return cudaLaunchKernelEx(&cfg, chain_persistent_kernel<16, 8>, p);
    } else if (N == 512) {
        cfg.dynamicSmemBytes = Tables<16, 4>::SMEM;
        return cudaLaunchKernelEx(&cfg, chain_persistent_kernel<16, 4>, p);
    } else if (T == 8) {
        cfg.dynamicSmemBytes = Tables<32, 8>::SMEM;
        return cudaLaunchKernelEx(&cfg, chain_persistent_kernel<32, 8>, p);
    }
    cfg.dynamicSmemBytes = Tables<32, 4>::SMEM;
    return cudaLaunchKernelEx(&cfg, chain_persistent_kernel<32, 4>, p);
}

// development aid (WRP_DEBUG=16): how many queue items found their dependency unmet when probed
void persistent_debug_counters(const int *ctrl, int *not_ready_a, int *not_ready_b)
{
    int v[3] = {0, 0, 0};
    cudaMemcpy(v, ctrl, sizeof v, cudaMemcpyDeviceToHost);
    *not_ready_a = v[1];
    *not_ready_b = v[2];
}

} // namespace wrp
