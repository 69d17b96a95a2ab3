// wrp_unified.cu — the fused chain for the default sector shape (M = 1024, N = 512) as ONE
// persistent kernel whose work item is a range tile PLUS eight Doppler rows (sm_100a).
//
// A sector has 64 C range tiles (8 columns x 1024 rows of one channel plane) and 512 C Doppler rows
// (range gates that survive stage 04) — exactly eight rows per tile.  Item i of the queue is
//     A: range tile  (i % TA) of sector  i / TA            (stages 01-02, as in wrp_persistent.cu)
//     B: row group   (i % TA) of sector  i / TA - lag      (stages 03-10, energy form)
// so every item carries the same work, the x2 hand-off lives in an L2-resident ring of `ring`
// sector slots, and — unlike the two-kind queue of wrp_persistent.cu, where a short Doppler block
// cannot hide the DRAM fetch of the range tile that follows it — every load is prefetched a whole
// item ahead: the tile into the warp's 8 KiB region of the tile buffer right after the warp's last
// read of it, the Doppler row (4 KiB, one per warp) into its own buffer right after the exchange
// barrier.  The next item is published by thread 0 before that barrier, so nobody ever spins for it.
//
// Shared memory per CTA (two CTAs per SM): tile 64 KiB + rows 32 KiB + tables 15.5 KiB = 111.5 KiB.
//
// Ordering across CTAs (counters in p.ctrl, as in wrp_persistent.cu):
//   a_done[s] += 1 per finished range tile of sector s   (red.release after the NEXT CTA barrier);
//   b_done[s] += 1 per row group of sector s whose rows have landed in shared memory;
//   item (sA, sub) may load its rows once a_done[sA - lag] == TA and its tile's ring slot is free
//   once b_done[sA - ring] == TA.  Both refer to strictly earlier queue items: no deadlock.
#include <cstdlib>
#include <cstring>

#include "wrp_chain_params.h"
#include "wrp_fft.cuh"
#include "wrp_internal.h"
#include "wrp_ptx.cuh"

namespace wrp {

namespace uni {
constexpr int N = 512, R1B = 16, T = 8, NW = 8, THREADS = 256, R = 32;
constexpr int PITCH = T * 8;        // bytes per range-tile row
constexpr int OFF_ROWS = 65536;     // Doppler rows, 4 KiB per warp
constexpr int WRC_ROW = 32 * 4 + 16; // wr(i)*c transposed [32 b][32 a] floats, rows padded by 16 B
constexpr int TWA_ROW = 32 * 8 + 16; // range inter-pass twiddles [32 b][32 ka] float2
constexpr int OFF_WRC = OFF_ROWS + 32768;
constexpr int OFF_TWA = OFF_WRC + 32 * WRC_ROW;
constexpr int OFF_WD = OFF_TWA + 32 * TWA_ROW;
constexpr int OFF_TL = OFF_WD + N * 4;
constexpr int SMEM = OFF_TL + 32 * 16;
static_assert(2 * (SMEM + 1024 + 256) <= 233472, "two CTAs per SM");

struct Item {
    int sa;   // sector of the range tile (the Doppler rows belong to sector sa - lag); < 0: queue empty
    int sub;  // tile / row-group index inside the sector
    int slot_a, slot_b; // x2 ring slots of sector sa and sector sa - lag
};
} // namespace uni

using uni::Item;


__device__ __forceinline__ Item uni_decode(int idx, const PersistParams &p, int ta)
{
    Item it;
    it.sa = idx / ta;
    it.sub = idx - it.sa * ta;
    it.slot_a = it.sa % p.ring;
    it.slot_b = it.sa >= p.lag ? (it.sa - p.lag) % p.ring : 0;
    return it;
}

// x2-ring row of warp `warp` in row group `sub`, and its (channel, gate):
//   pair groups (C >= 2): four gates x (hh, vv) — neighbouring warps hold hh and vv of one gate;
//   the rest: eight gates of vh (or of the only channel)
__device__ __forceinline__ const uint8_t *uni_row(const PersistParams &p, int slot, int sub, int pair_groups, int warp,
                                                  int &chn, int &gate)
{
    if (sub < pair_groups) {
        chn = warp & 1;
        gate = sub * 4 + (warp >> 1);
    } else {
        chn = p.C == 1 ? 0 : 2;
        gate = (sub - pair_groups) * 8 + warp;
    }
    return (const uint8_t *)p.x2 + (((size_t)slot * p.C + chn) * p.half_m + gate) * (size_t)(uni::N * 8);
}

// the warp's Doppler row of item `it` -> its 4 KiB row buffer (8 cp.async per lane, L2 hits)
__device__ __forceinline__ void uni_issue_row(const Item &it, const PersistParams &p, uint8_t *smem, int pair_groups,
                                              int warp, int lane)
{
    int chn, gate;
    const uint8_t *src = uni_row(p, it.slot_b, it.sub, pair_groups, warp, chn, gate) + lane * 16;
    uint8_t *dst = smem + uni::OFF_ROWS + warp * 4096 + lane * 16;
#ifdef WRP_UNI_ROWS_EVICT_FIRST
    // EXPERIMENT (not validated on a GPU): a ring row is dead once it has been read — let L2 drop it first
    const uint64_t pol = policy_evict_first();
    const uint32_t d = opaque_smem_addr(dst, p.zero);
#pragma unroll
    for (int k = 0; k < 8; ++k) cp_async16_evict_first(d + k * 512, src + k * 512, pol);
#else
#pragma unroll
    for (int k = 0; k < 8; ++k) cp_async16(dst + k * 512, src + k * 512);
#endif
}

// the warp's 8 KiB region (rows [128 warp, +128)) of the range tile of item `it`, then one arrival
// on the tile barrier
__device__ __forceinline__ void uni_issue_tile(const Item &it, const PersistParams &p, uint8_t *smem, uint64_t *bar,
                                               int warp, int lane)
{
    const int ch = it.sub >> 6, col_tile = it.sub & 63;
    const uint8_t *src = (const uint8_t *)p.iq +
                         ((size_t)(it.sa * p.C + ch) * 1024 + warp * 128 + (lane >> 2)) * (uni::N * 8) + col_tile * 64 +
                         (lane & 3) * 16;
    uint8_t *dst = smem + warp * 8192 + lane * 16;
    if (p.evict_first) {
        const uint64_t pol = policy_evict_first();
        const uint32_t d = opaque_smem_addr(dst, p.zero);
#pragma unroll
        for (int k = 0; k < 16; ++k) cp_async16_evict_first(d + k * 512, src + (size_t)k * 8 * (uni::N * 8), pol);
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) cp_async16(dst + k * 512, src + (size_t)k * 8 * (uni::N * 8));
    }
    cp_async_arrive(bar);
}

__global__ void __launch_bounds__(uni::THREADS, 2) chain_unified_kernel(const PersistParams p)
{
    using namespace uni;
    constexpr int SW = 128 / PITCH - 1; // row-swizzle mask of the in-place exchange (= 1)
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *const tile = smem;
    __shared__ __align__(8) uint64_t mbar; // the item's range tile has landed (one arrival per thread)
    __shared__ int4 s_item[2];             // published items, by item parity
    __shared__ int s_go[2];                // ... and whether their dependencies were met when probed
    __shared__ float p_row[2][NW];         // row powers of the current item, by item parity
#ifdef WRP_UNI_SPLIT_BARRIER
    // EXPERIMENT (not validated on a GPU; the default build is unaffected): the exchange rendezvous as a
    // split-phase mbarrier — arrive after the exchange stores, do the Doppler row's arithmetic, then wait
    __shared__ __align__(8) uint64_t xbar;
    uint32_t xphase = 0;
#endif

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TA = 64 * p.C;                      // tiles (= row groups) per sector
    const int TGT_A = TA;
    const int pair_groups = p.C >= 2 ? 128 : 0;   // row groups holding (hh, vv) pairs
    const int total = (p.S + p.lag) * TA;
    int *const a_done = p.ctrl + CTRL_A, *const b_done = p.ctrl + CTRL_A + p.smax;

    for (int i = tid; i < 32 * 32; i += THREADS) {
        *reinterpret_cast<float *>(smem + OFF_WRC + (i >> 5) * WRC_ROW + (i & 31) * 4) = __ldg(p.wrc_t + i);
        *reinterpret_cast<float2 *>(smem + OFF_TWA + (i >> 5) * TWA_ROW + (i & 31) * 8) = __ldg(p.tw_a + i);
    }
    for (int i = tid; i < N; i += THREADS) reinterpret_cast<float *>(smem + OFF_WD)[i] = __ldg(p.wd + i);
    if (tid < 32) {
        // t_m(l) = (-1)^l exp(-2 pi i l m / N), m = 1, 2: lane l's factor of clipped bin N/2 - m
        float s1, c1, s2, c2;
        sincospif(-2.f * (float)tid / (float)N, &s1, &c1);
        sincospif(-4.f * (float)tid / (float)N, &s2, &c2);
        const float sg = (tid & 1) ? -1.f : 1.f;
        *reinterpret_cast<float4 *>(smem + OFF_TL + tid * 16) = make_float4(sg * c1, sg * s1, sg * c2, sg * s2);
    }

    // dependencies of an item: tile -> ring slot free (b_done of sector sa - ring), rows -> every
    // range tile of sector sa - lag published (a_done)
    auto dep_a = [&](const Item &x) -> const int * {
        return (x.sa < p.S && x.sa >= p.ring) ? b_done + (x.sa - p.ring) : nullptr;
    };
    auto dep_b = [&](const Item &x) -> const int * {
        return (x.sa >= p.lag) ? a_done + (x.sa - p.lag) : nullptr;
    };

    // Items are dealt round-robin: item k of CTA x is queue index x + k * gridDim.x.  Every item
    // carries the same work, so a dynamic queue would balance nothing — and claiming two items ahead
    // (to hide the atomic) widens the window of started-but-unpublished tiles by two items per CTA,
    // which at lag 4 left 70 % of the items waiting for a dependency (ncu: the wait path's barrier).
    int claimed_next = 0; // thread 0: queue index of the item after the current one
    if (tid == 0) {
        mbar_init(&mbar, THREADS);
#ifdef WRP_UNI_SPLIT_BARRIER
        mbar_init(&xbar, THREADS);
#endif
        const int first = blockIdx.x;
        claimed_next = first + gridDim.x;
        Item f{-1, 0, 0, 0};
        if (first < total) {
            f = uni_decode(first, p, TA);
            if (const int *d = dep_a(f)) spin_until(d, TA); // nothing is held yet: blocking is safe
            if (const int *d = dep_b(f)) spin_until(d, TGT_A);
        }
        s_item[0] = make_int4(f.sa, f.sub, f.slot_a, f.slot_b);
        s_go[0] = 3;
    }
    __syncthreads();
    Item it{s_item[0].x, s_item[0].y, s_item[0].z, s_item[0].w};

    // every item's loads are two cp.async groups per thread, committed in this order: row, tile part
    auto issue_loads_row = [&](const Item &x) {
        // (one 4 KiB cp.async.bulk per warp instead of eight cp.async per lane: measured neutral, 279.6k vs 280.1k)
        if (x.sa >= p.lag) uni_issue_row(x, p, smem, pair_groups, warp, lane);
        cp_async_commit();
    };
    auto issue_loads_tile = [&](const Item &x) {
        if (x.sa < p.S) uni_issue_tile(x, p, smem, &mbar, warp, lane);
        cp_async_commit();
    };
    if (it.sa >= 0) {
        issue_loads_row(it);
        issue_loads_tile(it);
    }

    uint32_t phase = 0;
    int n = 0;        // index of the current item in this CTA's sequence
    int pending = -1; // sector of a finished range tile whose completion this CTA has not published yet
    bool rows_late = false; // the current item's rows were requested after its tile (dependency wait)

    while (it.sa >= 0) {
        const bool has_a = it.sa < p.S, has_b = it.sa >= p.lag;
        const int nslot = (n + 1) & 1;
        // Thread 0 works one item ahead: decode item n+1 (claimed during item n-1), probe its
        // dependencies, claim item n+2.  The round trips hide behind the Doppler row and the first
        // FFT pass; the result is published before this item's CTA barrier.
        Item cand{-1, 0, 0, 0};
        const int *pa = nullptr, *pb = nullptr;
        int va = 0, vb = 0, claimed_next2 = 0;
        if (tid == 0) {
            if (claimed_next < total) {
                cand = uni_decode(claimed_next, p, TA);
                pa = dep_a(cand);
                pb = dep_b(cand);
                if (pa) va = ld_relaxed(pa);
                if (pb) vb = ld_relaxed(pb);
            }
            claimed_next2 = claimed_next + gridDim.x;
        }
        auto publish_next = [&]() {
            if (tid == 0) {
                bool ok_a = !pa || va >= TA, ok_b = !pb || vb >= TGT_A;
                if (!ok_a) ok_a = ld_relaxed(pa) >= TA; // the early probe may be stale
                if (!ok_b) ok_b = ld_relaxed(pb) >= TGT_A;
                if ((p.debug & 16) && cand.sa >= 0) {
                    if (!ok_a) atomicAdd(p.ctrl + 1, 1); // ring slot still being read
                    if (!ok_b) atomicAdd(p.ctrl + 2, 1); // tiles of the rows' sector unpublished
                }
                s_item[nslot] = make_int4(cand.sa, cand.sub, cand.slot_a, cand.slot_b);
                s_go[nslot] = (ok_a ? 1 : 0) | (ok_b ? 2 : 0);
            }
        };
        // (prefetch.global.L2 of the item-after-next's tile, which round-robin dealing makes known two
        // items ahead, was measured 16-22 % SLOWER — 233k / 215k sectors/s with one / two prefetches per
        // 64-byte row segment against 276k without: the requests compete with the cp.async stream.)

        // ================= Doppler row of this warp: stages 03-08 in energy form =================
        // (see wrp_persistent.cu, DOP == 1, for the derivation)  P = N E - |Y_0|^2 - |Y_{N/2-1}|^2 - |Y_{N/2-2}|^2
        float pw = 0.f;
        float2 rv[R1B]; // the warp's Doppler row, lane l holds x[32 a + l]
        auto row_load = [&]() {
            // this thread's share of the row; the tile group, committed after it, may still be in flight
            if (rows_late) cp_async_wait_group<0>();
            else cp_async_wait_group<1>();
            __syncwarp();
            const uint8_t *row = smem + OFF_ROWS + warp * 4096;
            static_for<R1B>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                rv[a] = *reinterpret_cast<const float2 *>(row + (32 * a + lane) * 8);
            });
        };
        auto row_math = [&]() {
            const float4 tl = *reinterpret_cast<const float4 *>(smem + OFF_TL + lane * 16);
            float2 e2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            static_for<R1B>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                e2[a & 1] = cfma2(rv[a], rv[a], e2[a & 1]);
            });
            const float2 es = cadd(e2[0], e2[1]);
            float2 b0, b1, b2;
            dft_bins012<R1B>(rv, b0, b1, b2);
            const float2 y1 = cmul(b1, make_float2(tl.x, tl.y)), y2 = cmul(b2, make_float2(tl.z, tl.w));
            float r[7] = {es.x + es.y, b0.x, b0.y, y1.x, y1.y, y2.x, y2.y};
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < 7; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
            }
            float removed = r[1] * r[1];
#pragma unroll
            for (int k = 2; k < 7; ++k) removed = fmaf(r[k], r[k], removed);
            pw = fmaxf(fmaf((float)N, r[0], -removed), 0.f) * p.taps_sum; // stages 05-08: x sum of the taps
            if (lane == 0) p_row[n & 1][warp] = pw;
        };
#ifndef WRP_UNI_SPLIT_BARRIER
        if (has_b) {
            row_load();
            row_math();
        }
#endif

        // ================= range tile, first pass =================
        const int c = tid % T, b = tid / T;
        const int ch = it.sub >> 6, col = (it.sub & 63) * T + c;
        float2 v[R];
        if (has_a) {
            mbar_wait(&mbar, phase);
            phase ^= 1;
            {
                const uint8_t *src = tile + b * PITCH + c * 8;
                static_for<R>([&](auto ai) {
                    constexpr int a = decltype(ai)::value;
                    v[brev<R>(a)] = *reinterpret_cast<const float2 *>(src + a * (R * PITCH));
                });
            }
            {
                // stage 01 (x *= wr(i)*c*wd(j), rpv2.cu:86-91) fused into the first butterfly stage:
                // the span-1 partners of the bit-reversed network are rows a and a + 16
                const float wdj = reinterpret_cast<const float *>(smem + OFF_WD)[col];
                const float2 m2 = make_float2(-2.f, -2.f);
                const float4 *w4 = reinterpret_cast<const float4 *>(smem + OFF_WRC + b * WRC_ROW);
                static_for<R / 8>([&](auto qi) { // rows 4q .. 4q+3 and their partners 16 + 4q ..
                    constexpr int q = decltype(qi)::value;
                    const float4 wlo = w4[q], whi = w4[q + R / 8];
                    const float lo[4] = {wlo.x, wlo.y, wlo.z, wlo.w}, hi[4] = {whi.x, whi.y, whi.z, whi.w};
                    static_for<4>([&](auto ei) {
                        constexpr int e = decltype(ei)::value;
                        constexpr int sa = brev<R>(4 * q + e); // even slot; partner row a + R/2 sits in sa + 1
                        static_assert(brev<R>(4 * q + e + R / 2) == sa + 1, "span-1 partner");
                        const float wl = lo[e] * wdj, wh = hi[e] * wdj;
                        const float2 t = cmul2(v[sa + 1], make_float2(wh, wh));
                        const float2 s2 = cfma2(v[sa], make_float2(wl, wl), t); // A*wl + B*wh
                        v[sa + 1] = cfma2(t, m2, s2);                            // A*wl - B*wh
                        v[sa] = s2;
                    });
                });
                fft_dit_after_stage1<R, -1>(v);
            }
            publish_next();
            __syncwarp();
            {
                // Z[ka][b] goes to row 32 ka + (b ^ (ka & SW)): the warp keeps its own row footprint (in
                // place); the twiddle reads run two steps ahead of the exchange stores
                const float4 *t4 = reinterpret_cast<const float4 *>(smem + OFF_TWA + b * TWA_ROW);
                uint8_t *d_sw[SW + 1];
#pragma unroll
                for (int sx = 0; sx <= SW; ++sx) d_sw[sx] = tile + (b ^ sx) * PITCH + c * 8;
                float4 wq[3] = {t4[0], t4[1], t4[2]};
                static_for<R / 2>([&](auto qi) {
                    constexpr int q = decltype(qi)::value;
                    const float4 w = wq[q % 3];
                    if constexpr (q + 3 < R / 2) wq[q % 3] = t4[q + 3];
                    const float2 y0 = q == 0 ? v[0] : cmul(v[2 * q], make_float2(w.x, w.y));
                    const float2 y1 = cmul(v[2 * q + 1], make_float2(w.z, w.w));
                    *reinterpret_cast<float2 *>(d_sw[(2 * q) & SW] + (2 * q) * (R * PITCH)) = y0;
                    *reinterpret_cast<float2 *>(d_sw[(2 * q + 1) & SW] + (2 * q + 1) * (R * PITCH)) = y1;
                });
            }
        } else {
            publish_next();
        }
#ifdef WRP_UNI_SPLIT_BARRIER
        if (has_b) row_load(); // in registers before the arrival, so b_done may follow the wait
        mbar_arrive(&xbar);
        if (has_b) row_math(); // overlaps the wait for the slower warps
        mbar_wait(&xbar, xphase);
        xphase ^= 1;
#else
        __syncthreads(); // the exchange — and the one rendezvous of the item
#endif

        auto write_products = [&](int sb) {
            if (lane == 0) {
                int chn, gate;
                (void)uni_row(p, it.slot_b, it.sub, pair_groups, warp, chn, gate);
                if (p.power) p.power[((size_t)sb * p.C + chn) * p.half_m + gate] = pw;
                const bool pair = it.sub < pair_groups;
                if ((pair && !(warp & 1)) || (!pair && p.C == 1)) {
                    // stages 09/10 (rpv2.cu:199-213); vv of the gate sits in the neighbouring warp
                    const float rg = (float)gate * p.range_res;
                    const float z = rg * rg * p.calib * pw;
                    reinterpret_cast<float2 *>(p.out)[(size_t)sb * p.half_m + gate] =
                        make_float2(10.f * log10f(z), pair ? 10.f * (log10f(pw) - log10f(p_row[n & 1][warp + 1])) : 0.f);
                }
            }
        };
        // ---- right after the barrier: publications, products, the next item's Doppler row ----
        if (pending >= 0 && tid == THREADS - 32) red_release_add(a_done + pending); // every warp's x2 stores precede the barrier
        pending = -1;
        if (has_b) {
            const int sb = it.sa - p.lag;
            // every warp's row has landed: the ring rows are free.  (Counting per warp at the top of the
            // item instead — eight atomics on one address per item — was measured 11 % slower: the
            // counter becomes an L2 hot spot and the dependency waits double.)
            if (tid == THREADS - 64) atomicAdd(b_done + sb, 1);
#ifndef WRP_UNI_SPLIT_BARRIER
            write_products(sb);
#endif
        }
        const Item nit{s_item[nslot].x, s_item[nslot].y, s_item[nslot].z, s_item[nslot].w};
        const int go_bits = nit.sa >= 0 ? s_go[nslot] : 0;
        const bool go_a = go_bits & 1, go_b = go_bits & 2; // tile / rows may be fetched now
        if (go_b) issue_loads_row(nit);
#ifdef WRP_UNI_SPLIT_BARRIER
        if (has_b) { // p_row of the neighbouring warp was written after its arrival: meet it first
            bar_sync(1 + (warp >> 1), 64);
            write_products(it.sa - p.lag);
        }
#endif

        // ================= range tile, second pass =================
        if (has_a) {
            const int ka = b; // rows 32 ka + .. of warp w (ka = 4 w ..) are its own 8 KiB region
            {
                const uint8_t *s_sw[SW + 1];
#pragma unroll
                for (int sx = 0; sx <= SW; ++sx) s_sw[sx] = tile + ka * (R * PITCH) + c * 8 + ((ka ^ sx) & SW) * PITCH;
                static_for<R>([&](auto bi) {
                    constexpr int bb = decltype(bi)::value;
                    v[brev<R>(bb)] = *reinterpret_cast<const float2 *>(s_sw[bb & SW] + (bb & ~SW) * PITCH);
                });
            }
            __syncwarp();
            if (go_a) issue_loads_tile(nit); // the warp's region is in registers: fetch its share of the next tile
            fft_dit<R, -1>(v);
            {
                float2 *out = p.x2 + (((size_t)it.slot_a * p.C + ch) * p.half_m + ka) * (size_t)N + col;
                static_for<R / 2>([&](auto ki) { // rows k = ka + 32 kb < M/2
                    constexpr int kb = decltype(ki)::value;
#ifdef WRP_UNI_X2_EVICT_LAST
                    // EXPERIMENT (not validated on a GPU): keep the hand-off in L2 against the input stream
                    st_global_hint(out + (size_t)(R * kb) * N, v[kb], policy_evict_last());
#else
                    out[(size_t)(R * kb) * N] = v[kb];
#endif
                });
            }
            // published after the next CTA barrier, when these stores have drained (a red.release per
            // warp right here was measured 17 % slower: the warp idles until its stores are visible)
            pending = it.sa;
        } else if (go_a) {
            issue_loads_tile(nit);
        }

        rows_late = false;
        if (nit.sa < 0 || !go_a || !go_b) {
            // leaving, or a dependency of the next item was unmet when probed (it may be a tile this very
            // CTA still holds unpublished): publish, then thread 0 waits for the counter(s) and the loads
            // that were held back go out.  A tile whose ring slot was free has already been requested.
            __syncthreads();
            if (pending >= 0 && tid == THREADS - 32) red_release_add(a_done + pending);
            pending = -1;
            if (nit.sa >= 0) {
                if (tid == 0) {
                    if (!go_a)
                        if (const int *d = dep_a(nit)) spin_until(d, TA);
                    if (!go_b)
                        if (const int *d = dep_b(nit)) spin_until(d, TGT_A);
                }
                __syncthreads();
                if (!go_b) issue_loads_row(nit);
                if (!go_a) issue_loads_tile(nit);
                rows_late = go_a && !go_b; // the row group was committed after the tile group
            }
        }
        it = nit;
        ++n;
        claimed_next = claimed_next2;
    }
}

// ---- host side ---------------------------------------------------------------------------------
bool unified_supported(int M, int N) { return M == 1024 && N == 512; }

cudaError_t unified_setup()
{
    return cudaFuncSetAttribute(chain_unified_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, uni::SMEM);
}

cudaError_t launch_unified(PersistParams p, int sm_count, cudaStream_t st)
{
    if (p.lag > p.S) p.lag = p.S; // short batches: no item without a tile and without rows
    const int total = (p.S + p.lag) * 64 * p.C;
    int grid = 2 * sm_count;
    if (grid > total) grid = total;
    chain_unified_kernel<<<grid, uni::THREADS, uni::SMEM, st>>>(p);
    return cudaGetLastError();
}

} // namespace wrp
