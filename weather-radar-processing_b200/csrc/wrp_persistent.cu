// wrp_persistent.cu — the fused chain as ONE persistent kernel per batch (sm_100a): the two-kind
// work queue with an L2-resident range -> Doppler hand-off ring.  The product path is the hand-off-free
// streaming kernel (wrp_stream.cu); this file carries the literal Doppler transform
// (wrp_config.doppler_form = WRP_DOPPLER_FFT) and stays selectable (chain_impl = WRP_CHAIN_QUEUE).
//
// Work items, handed out in queue order by an atomic counter:
//   A(s, tile)  range tile: T (= 8) adjacent columns x 1024 rows of one (sector, channel) plane.
//               cp.async (16 B, L1 bypass) -> window folded into the first butterfly stage ->
//               radix-32 x radix-32 column FFT with an in-place shared-memory exchange -> rows
//               k < M/2 to the x2 ring (stages 01-02; replaces __apply_hamming + the strided cuFFT
//               plan, rpv2.cu:86-91, 318-333, 418-428).
//   B(s, block) Doppler block: 16 rows (8 gates x hh,vv or 16 gates of vh) of the x2 ring ->
//               mean removal, inverse DFT, shift, clip, |.|^2, moving-average power, ZdB/ZDR
//               (stages 03-10; replaces rpv2.cu:93-213, 434-566).  Rows are warp-private.
//               DOP = 1 (default): stages 03-08 in energy form — row energy minus the DC bin and the
//               two clipped bins (Parseval), one pass over the row; DOP = 0 (WRP_DOPPLER=fft): the
//               literal two-pass transform.
// The queue interleaves A(t) with B(t - lag), so the x2 hand-off lives in a small ring
// (ring x 6 MiB) that never leaves L2, there is no kernel boundary between the phases, and
// DRAM-heavy A items overlap compute-heavy B items on every SM.
//
// Synchronisation inside a CTA.  The 64 KiB tile buffer is cut into one 8 KiB region per warp;
// the last shared-memory reads of an item (second FFT pass) touch only the warp's own region for
// BOTH item kinds, and the next item's bytes for that region are fetched by that warp.  So a warp
// prefetches the moment it has pulled its own operands — no buffer-release barrier — and loads
// overlap the second pass and the epilogue.  An mbarrier (one arrival per thread, fired by
// cp.async completion) tells everybody when the whole next tile has landed.  The only CTA barrier
// left is the range tile's exchange between its two FFT passes; Doppler blocks have none.
// Thread 0 claims the next queue slot at the top of an item, decodes and probes it mid-item and
// publishes it through shared memory (sequence counters, no barrier).
//
// Cross-CTA ordering: per-sector completion counters.  A warp publishes its share of a finished
// range tile with red.release.gpu later on (after the first pass of its next item, when the
// stores have long drained); Doppler blocks are only loaded once their sector's count is full
// (ld.relaxed.gpu probe; data is then read with cp.async.cg straight from L2).  An item only waits
// for items earlier in the queue, and nobody spins while holding unpublished work: no deadlock.
#include "wrp_fft.cuh"
#include "wrp_internal.h"
#include "wrp_ptx.cuh"
#include "wrp_chain_params.h"

namespace wrp {

// shared-memory copies of the small tables; rows are padded by 16 B so that the 128-bit reads of
// lanes holding different rows hit different banks
constexpr int WRC_ROW = 32 * 8 + 16; // wr(i)*c transposed [32][32], each value stored twice (w, w) for FMUL2
constexpr int TWA_ROW = 32 * 8 + 16; // range inter-pass twiddles [32][32] float2
constexpr int TAB_WRC = 32 * WRC_ROW;
constexpr int TAB_TWA = 32 * TWA_ROW;
// T = columns per range tile, Q = M / 1024; T*Q warps per CTA, the tile buffer holds 8 KiB per warp
template <int R1B, int T, int Q> struct Tables {
    static constexpr int NW = T * Q;
    static constexpr int TILE_BYTES = NW * 8192;
    static constexpr int N = 32 * R1B;
    static constexpr int TWB_ROW = R1B * 8 + 16; // Doppler inter-pass twiddles [32][R1B] float2
    static constexpr int WD = N * 4;             // Doppler window
    static constexpr int TWB = 32 * TWB_ROW;
    // Q = 1: transposed (w, w) window of the fused first stage; Q = 4: wr(i)*c [4096] and the
    // radix-4 pre-pass twiddle W^r [1024] (W^2r, W^3r are formed by multiplication)
    static constexpr int OFF_WRC = TILE_BYTES;
    static constexpr int OFF_W4 = TILE_BYTES;
    static constexpr int OFF_TW4 = OFF_W4 + 4096 * 4;
    static constexpr int OFF_TWA = Q == 1 ? OFF_WRC + TAB_WRC : OFF_TW4 + 1024 * 8;
    static constexpr int OFF_TWB = OFF_TWA + TAB_TWA;
    static constexpr int OFF_WD = OFF_TWB + TWB;
    // energy form of the Doppler stage: per-lane factors of the two clipped bins, [32] float4
    static constexpr int OFF_TL = OFF_WD + WD;
    static constexpr int SMEM = OFF_TL + 32 * 16;
};

struct Item {
    int kind; // 0 = range tile, 1 = Doppler block, < 0 = queue empty
    int sector;
    int sub;
    int slot; // x2 ring slot = sector % ring (a runtime division: done once, by the decoding thread)
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
        it.slot = it.sector % p.ring;
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
        it.slot = it.sector % p.ring;
        return it;
    }
    idx -= p.n2 * per;
    it.kind = 1;
    const int t = idx / TB;
    it.sector = p.b3_first + t;
    it.sub = idx - t * TB;
    it.slot = it.sector % p.ring;
    return it;
}

// dependency of an item: the counter it must see reach `target` before its copy may start
//   range tile of sector s >= ring: WAR on ring slot — the Doppler blocks of sector s - ring must
//                                   have pulled their rows;
//   Doppler block of sector s:      every range tile of sector s has been published.
template <int T> __device__ __forceinline__ const int *item_dep(const Item &it, const PersistParams &p, int &target)
{
    if (it.kind == 0) {
        target = p.blocks_b;
        return it.sector >= p.ring ? p.ctrl + CTRL_A + p.smax + (it.sector - p.ring) : nullptr;
    }
    target = p.tiles_a;
    return p.ctrl + CTRL_A + it.sector;
}

// x2-ring address of row (warp, rr) of a Doppler block, and its (channel, gate)
template <int N, int NW> // NW = warps per CTA
__device__ __forceinline__ const uint8_t *doppler_row(const Item &it, const PersistParams &p, int warp, int rr,
                                                      int &chn, int &gate)
{
    constexpr int ROWS_B = NW * 1024 / N, RPW = ROWS_B / NW;
    if (it.sub < p.pair_blocks) {
        chn = RPW == 2 ? rr : (warp & 1);
        gate = it.sub * (ROWS_B / 2) + (RPW == 2 ? warp : (warp >> 1));
    } else {
        chn = p.C == 1 ? 0 : 2;
        gate = (it.sub - p.pair_blocks) * ROWS_B + warp * RPW + rr;
    }
    return (const uint8_t *)p.x2 + (((size_t)it.slot * p.C + chn) * p.half_m + gate) * (size_t)(N * 8);
}
// The ring rows are scratch: once a Doppler block has them in shared memory their L2 lines are
// dropped, so the dirty lines are never written back to DRAM
__device__ __forceinline__ void discard_l2_line(const void *p)
{
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}

// The executing warp fetches its own 8 KiB region of the item's tile (16 cp.async per lane).
//   range tile:    rows [1024/T * warp, +1024/T) of the T-column tile (16-byte chunks, T/2 per row)
//   Doppler block: the warp's RPW rows of the x2 ring, contiguous N*8 bytes each
template <int N, int T, int Q>
__device__ __forceinline__ void issue_warp_load(const Item &it, const PersistParams &p, uint8_t *tile, uint64_t *bar,
                                                int warp, int lane, uint32_t after_and_zero = 0)
{
    // after_and_zero = (a value derived from the warp's LAST read of the region) & p.zero: always 0, but it makes the
    // copies below data-dependent on the completion of those reads — cp.async data lands asynchronously, and being
    // ordered behind the mere ISSUE of the loads is not enough (wrp_stream.cu, tma_load_2d, has the full story)
    uint8_t *dst = tile + warp * 8192 + lane * 16 + after_and_zero;
    if (it.kind == 0) {
        constexpr int tiles_per_plane = N / T, CPR = T / 2, RPWARP = 1024 / T;
        const int ch = it.sub / tiles_per_plane, col_tile = it.sub - ch * tiles_per_plane;
        const uint8_t *src = (const uint8_t *)p.iq +
                             ((size_t)(it.sector * p.C + ch) * (1024 * Q) + warp * RPWARP + lane / CPR) * (N * 8) +
                             col_tile * (T * 8) + (lane % CPR) * 16;
        // (N = 512 instances only: in the spilling N = 1024 instances ptxas picks an odd descriptor
        // register for the hinted form — see cp_async16_evict_first)
        if (p.evict_first) {
            const uint64_t pol = policy_evict_first();
            const uint32_t d = opaque_smem_addr(dst, p.zero);
#pragma unroll
            for (int k = 0; k < 16; ++k)
                cp_async16_evict_first(d + k * 512, src + (size_t)k * (32 / CPR) * (N * 8), pol);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) cp_async16(dst + k * 512, src + (size_t)k * (32 / CPR) * (N * 8));
        }
    } else {
        constexpr int NW = T * Q, ROWS_B = NW * 1024 / N, RPW = ROWS_B / NW;
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            int chn, gate;
            const uint8_t *src = doppler_row<N, NW>(it, p, warp, rr, chn, gate) + lane * 16;
#pragma unroll
            for (int k = 0; k < 16 / RPW; ++k) cp_async16(dst + rr * (N * 8) + k * 512, src + k * 512);
        }
    }
    cp_async_arrive(bar);
}

// ---- the kernel ------------------------------------------------------------------------------
// Doppler length N = 32 * R1B; T columns per range tile; range length M = 1024 * Q (Q = 1 or 4).
// Q = 4: a radix-4 decimation-in-frequency pre-pass (window folded in) turns the T columns x 4096
// rows into 4T independent 1024-point columns, one per warp, whose outputs k' < 512 are the rows
// 4 k' + k0 < M/2 of the 4096-point transform — so the pruning and both 32 x 32 passes are shared.
template <int R1B, int T, int Q, int DOP>
__global__ void __launch_bounds__(32 * T * Q, 16 / (T * Q))
    chain_persistent_kernel(const PersistParams p)
{
    constexpr int R = 32; // (sub-)range FFT 32 x 32
    constexpr int N = 32 * R1B;
    constexpr int NW = T * Q; // warps
    constexpr int THREADS = 32 * NW;
    using Tab = Tables<R1B, T, Q>;
    constexpr int PITCH = T * 8;          // bytes per range-tile row
    constexpr int SW = 128 / PITCH - 1;   // row-swizzle mask of the in-place exchange
    constexpr int ROWS_B = NW * 1024 / N; // Doppler rows per block
    constexpr int RPW = ROWS_B / NW;      // rows per warp
    static_assert(Q == 1 || Q == 4, "range length 1024 or 4096");
    static_assert(RPW == 1 || RPW == 2, "Doppler rows per warp");
    extern __shared__ __align__(1024) uint8_t tile[];
    __shared__ __align__(8) uint64_t mbar; // next tile has landed: one arrival per thread, fired by its cp.asyncs
    __shared__ int4 s_item[2];             // published items (kind, sector, sub, -)
    __shared__ volatile int s_seq;         // items published so far
    __shared__ volatile int s_go;          // items whose dependency is known to be met
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
        if constexpr (Q == 1) {
            for (int i = tid; i < 32 * 32; i += THREADS) { // window values duplicated into (w, w) pairs
                const float w = __ldg(p.wrc_t + i);
                *reinterpret_cast<float2 *>(tile + Tab::OFF_WRC + (i >> 5) * WRC_ROW + (i & 31) * 8) =
                    make_float2(w, w);
            }
        } else {
            copy_rows(Tab::OFF_W4, p.wr4, 1, 4096 * 4, 4096 * 4);
            copy_rows(Tab::OFF_TW4, p.tw4, 1, 1024 * 8, 1024 * 8);
        }
        copy_rows(Tab::OFF_TWA, p.tw_a, 32, 32 * 8, TWA_ROW);
        copy_rows(Tab::OFF_TWB, p.tw_b, 32, R1B * 8, Tab::TWB_ROW);
        copy_rows(Tab::OFF_WD, p.wd, 1, Tab::WD, Tab::WD);
        if (DOP == 1 && tid < 32) {
            // t_m(l) = (-1)^l exp(-2 pi i l m / N), m = 1, 2: lane l's factor of clipped bin N/2 - m
            float s1, c1, s2, c2;
            sincospif(-2.f * (float)tid / (float)N, &s1, &c1);
            sincospif(-4.f * (float)tid / (float)N, &s2, &c2);
            const float sg = (tid & 1) ? -1.f : 1.f;
            *reinterpret_cast<float4 *>(tile + Tab::OFF_TL + tid * 16) = make_float4(sg * c1, sg * s1, sg * c2, sg * s2);
        }
    }
    int claimed_next = 0; // thread 0: queue index of the item after the current one (claimed one item ahead)
    if (tid == 0) {
        mbar_init(&mbar, THREADS);
        const int first = atomicAdd(p.ctrl, 1);
        claimed_next = atomicAdd(p.ctrl, 1);
        Item f{-1, 0, 0, 0};
        if (first < p.total_items) {
            f = decode_item(first, p);
            int target;
            const int *dep = item_dep<T>(f, p, target);
            if (dep) spin_until(dep, target); // nothing is held yet: blocking is safe
        }
        s_item[0] = make_int4(f.kind, f.sector, f.sub, f.slot);
        s_seq = 1;
        s_go = 1;
    }
    __syncthreads();
    Item it{s_item[0].x, s_item[0].y, s_item[0].z, s_item[0].w};
    if (it.kind >= 0) issue_warp_load<N, T, Q>(it, p, tile, &mbar, warp, lane);

    uint32_t phase = 0;
    int n = 0;        // index of the current item in this CTA's sequence
    int pending = -1;   // sector of a finished range tile whose completion this CTA has not published yet
    int pending_b = -1; // sector of a Doppler block whose ring rows were consumed (and discarded) but not yet reported
    // One release per item per CTA without a blocking CTA barrier: every warp but the last only
    // announces (bar.arrive) that it is past the stores / discards in question; the last warp waits
    // for the announcements (bar.sync) and its lane 0 publishes (warp 0 carries the queue work and
    // must not be held up).  Both variables are CTA-uniform (all warps walk the same item sequence).
    auto release_pending = [&]() {
        if (pending >= 0) red_release_add(p.ctrl + CTRL_A + pending);
        if (pending_b >= 0) red_release_add(p.ctrl + CTRL_A + p.smax + pending_b);
    };
    auto publish_pending = [&]() {
        if (pending >= 0 || pending_b >= 0) {
            if (warp == NW - 1) {
                bar_sync(1, THREADS);
                if (lane == 0) release_pending();
            } else {
                bar_arrive(1, THREADS);
            }
            pending = -1;
            pending_b = -1;
        }
    };
    // Same publication right after a CTA-wide barrier of the current item: the barrier already is
    // the rendezvous (every warp's stores of the previous item precede it), so nobody waits
    auto publish_pending_after_cta_barrier = [&]() {
        if (pending >= 0 || pending_b >= 0) {
            if (tid == THREADS - 32) release_pending();
            pending = -1;
            pending_b = -1;
        }
    };

    while (it.kind >= 0) {
        const int nslot = (n + 1) & 1;
        Item nit{-1, 0, 0, 0};
        bool loaded = false;
        // Thread 0 works one item ahead: the queue slot of item n+1 was claimed during item n-1, so
        // it can be decoded and its dependency probed right now; the claim for item n+2 goes out at
        // the same time.  Both round trips hide behind this item's first pass; the results are
        // published (shared memory, no barrier) right after it.
        Item cand{-1, 0, 0, 0};
        int probe = 0, probe_target = 0, claimed_next2 = 0;
        const int *probe_dep = nullptr;
        if (tid == 0) {
            if (claimed_next < p.total_items) {
                cand = decode_item(claimed_next, p);
                probe_dep = item_dep<T>(cand, p, probe_target);
                probe = probe_dep ? ld_relaxed(probe_dep) : (probe_target = 0);
            }
            claimed_next2 = atomicAdd(p.ctrl, 1);
        }
        auto publish_next = [&]() {
            if (tid == 0) {
                bool ready = probe >= probe_target;
                if (!ready) ready = ld_relaxed(probe_dep) >= probe_target; // the early probe may be stale
                if ((p.debug & 16) && cand.kind >= 0 && !ready) atomicAdd(p.ctrl + 1 + cand.kind, 1);
                s_item[nslot] = make_int4(cand.kind, cand.sector, cand.sub, cand.slot);
                if (ready) {
                    fence_acquire_gpu(); // one fence per item from one thread
                    s_go = n + 2;        // before s_seq: whoever sees the item also sees that it may load
                }
                __threadfence_block();
                s_seq = n + 2;
            }
        };
        // a warp that has pulled its last operand out of its region learns the next item and, if
        // that item's dependency is already met, starts fetching its own region of it
        auto prefetch_next = [&](uint32_t after_and_zero) {
            while (s_seq < n + 2) {
            }
            {
                const int4 pub = s_item[nslot];
                nit = Item{pub.x, pub.y, pub.z, pub.w};
            }
            if (nit.kind >= 0 && s_go >= n + 2) {
                issue_warp_load<N, T, Q>(nit, p, tile, &mbar, warp, lane, after_and_zero);
                loaded = true;
            }
        };

        mbar_wait(&mbar, phase);
        phase ^= 1;

        if (it.kind == 0) {
            // ================= range tile =================
            // Q = 4: thread group `sub` (32 T threads) owns the 1024-row sub-tile k0 = sub
            const int sub = Q == 1 ? 0 : tid / (32 * T), tl = Q == 1 ? tid : tid % (32 * T);
            const int c = tl % T, b = tl / T;
            constexpr int tiles_per_plane = N / T;
            const int ch = it.sub / tiles_per_plane, col0 = (it.sub - ch * tiles_per_plane) * T, col = col0 + c;
            uint8_t *stile = tile + sub * (1024 * PITCH);
            if constexpr (Q == 4) {
                // stage 01 + radix-4 DIF step over rows r, r + 1024, r + 2048, r + 3072, in place:
                //   y_k0[r] = W_4096^(r k0) * sum_q (-i)^(q k0) ham(r + 1024 q) x[r + 1024 q]
                // one unit = one row r x two adjacent columns (16-byte accesses)
                constexpr int UNITS = 1024 * T / 2;
                static_assert(THREADS % (T / 2) == 0, "column pair fixed per thread");
                const int cp = tid % (T / 2);
                const float2 wdp = *reinterpret_cast<const float2 *>(tile + Tab::OFF_WD + (col0 + 2 * cp) * 4);
                const float *wr4 = reinterpret_cast<const float *>(tile + Tab::OFF_W4);
                const float2 *tw4 = reinterpret_cast<const float2 *>(tile + Tab::OFF_TW4);
#pragma unroll 2
                for (int u = tid; u < UNITS; u += THREADS) {
                    const int r = u / (T / 2);
                    uint8_t *ptr = tile + r * PITCH + cp * 16;
                    float4 x[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) x[q] = *reinterpret_cast<const float4 *>(ptr + q * (1024 * PITCH));
                    float2 e[4], o[4]; // even / odd column of the pair
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float w = wr4[r + 1024 * q];
                        const float wa = w * wdp.x, wb = w * wdp.y;
                        e[q] = cmul2(make_float2(x[q].x, x[q].y), make_float2(wa, wa));
                        o[q] = cmul2(make_float2(x[q].z, x[q].w), make_float2(wb, wb));
                    }
                    const float2 w1 = tw4[r], w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                    auto radix4 = [&](float2 (&z)[4]) {
                        const float2 t0 = cadd(z[0], z[2]), t1 = csub(z[0], z[2]);
                        const float2 t2 = cadd(z[1], z[3]), d = csub(z[1], z[3]);
                        const float2 t3 = make_float2(d.y, -d.x); // -i (x1 - x3)
                        z[0] = cadd(t0, t2);
                        z[1] = cmul(cadd(t1, t3), w1);
                        z[2] = cmul(csub(t0, t2), w2);
                        z[3] = cmul(csub(t1, t3), w3);
                    };
                    radix4(e);
                    radix4(o);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4 *>(ptr + q * (1024 * PITCH)) = make_float4(e[q].x, e[q].y, o[q].x, o[q].y);
                }
                __syncthreads();
                publish_pending_after_cta_barrier();
            }
            float2 v[R];
            {
                const uint8_t *src = stile + b * PITCH + c * 8;
                static_for<R>([&](auto ai) {
                    constexpr int a = decltype(ai)::value;
                    v[brev<R>(a)] = *reinterpret_cast<const float2 *>(src + a * (R * PITCH));
                });
            }
            if constexpr (Q == 1) {
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
                fft_dit_after_stage1<R, -1>(v);
            } else {
                fft_dit<R, -1>(v);
            }
            publish_next();
            __syncwarp();
            {
                // Z[ka][b] goes to row 32 ka + (b ^ (ka & SW)): the warp keeps its own row footprint
                // (in place), and the 128/PITCH values of ka met by one shared-memory wavefront of
                // pass 2 fall into different PITCH-byte slices of a 128-byte bank line
                const float4 *t4 = reinterpret_cast<const float4 *>(tile + Tab::OFF_TWA + b * TWA_ROW);
                uint8_t *d_sw[SW + 1];
#pragma unroll
                for (int sx = 0; sx <= SW; ++sx) d_sw[sx] = stile + (b ^ sx) * PITCH + c * 8;
                // the twiddle reads run two steps ahead of the exchange stores: both go to the same
                // shared-memory array, so the compiler will not hoist a load above a store by itself
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
            // the exchange: the one barrier of a 1024-point column group (the whole CTA for Q = 1, the
            // 32 T threads of a sub-tile for Q = 4)
            if constexpr (Q == 1) {
                __syncthreads();
                publish_pending_after_cta_barrier();
            } else { // constant barrier ids, so that only 6 of the 16 are reserved
                if (sub == 0) bar_sync_id<2>(32 * T);
                else if (sub == 1) bar_sync_id<3>(32 * T);
                else if (sub == 2) bar_sync_id<4>(32 * T);
                else bar_sync_id<5>(32 * T);
            }
            const int ka = b; // rows 32 ka + .. of warp w (ka = 32/T * w ..) are its own 8 KiB region
            {
                // row 32 ka + (bb ^ (ka & SW)): SW + 1 swizzled bases, everything else is an immediate
                const uint8_t *s_sw[SW + 1];
#pragma unroll
                for (int sx = 0; sx <= SW; ++sx) s_sw[sx] = stile + ka * (R * PITCH) + c * 8 + ((ka ^ sx) & SW) * PITCH;
                static_for<R>([&](auto bi) {
                    constexpr int bb = decltype(bi)::value;
                    v[brev<R>(bb)] = *reinterpret_cast<const float2 *>(s_sw[bb & SW] + (bb & ~SW) * PITCH);
                });
            }
            __syncwarp();
            dit_stage<R, 1, -1>(v); // consumes every loaded value: the prefetch below follows the completion of the reads
            prefetch_next(__float_as_uint(v[R - 1].x) & (uint32_t)p.zero);
            fft_dit_after_stage1<R, -1>(v);
            {
                // sub-transform output k' = ka + 32 kb < 512 is row Q k' + sub of the M-point transform
                float2 *out =
                    p.x2 + (((size_t)it.slot * p.C + ch) * p.half_m + Q * ka + sub) * (size_t)N + col;
                static_for<R / 2>([&](auto ki) { // rows k < M/2
                    constexpr int kb = decltype(ki)::value;
                    out[(size_t)(Q * R * kb) * N] = v[kb];
                });
            }
            pending = it.sector; // published later (release), when these stores have drained
        } else {
            // ================= Doppler block =================
            // the ring rows are in shared memory now: drop their L2 lines (never written back), then
            // tell the range tiles that will reuse the ring slot — after every warp's discards
            if (p.discard) {
#pragma unroll
                for (int rr = 0; rr < RPW; ++rr) {
                    int chn_, gate_;
                    const uint8_t *row_g = doppler_row<N, NW>(it, p, warp, rr, chn_, gate_);
#pragma unroll
                    for (int k = 0; k < (N * 8) / (128 * 32); ++k) discard_l2_line(row_g + (k * 32 + lane) * 128);
                }
            }
            if (!p.discard) {
                if (tid == 0) red_release_add(p.ctrl + CTRL_A + p.smax + it.sector);
            } else {
                pending_b = it.sector; // reported after this item's first pass, when the discards have drained
            }
            const bool pair = it.sub < p.pair_blocks;
            uint8_t *region = tile + warp * 8192; // rows (warp, rr) live at region + rr * N*8
            const int rsel = RPW == 2 ? (lane >> 4) : 0; // which of the warp's rows this lane reports
            const int ka = RPW == 2 ? (lane & 15) : lane;
            float pw;
            if constexpr (DOP == 1) {
                // ---- energy form of stages 03-08 (Parseval) ----
                // The row power is the sum of |Y_b|^2 over every Doppler bin b except the DC bin (zeroed
                // by the mean removal, rpv2.cu:93-130) and the two clipped bins N/2-1, N/2-2 (columns
                // N-1, N-2 after the shift, rpv2.cu:137-148).  With Y the un-normalised transform,
                //   sum_b |Y_b|^2 = N sum_j |x_j|^2,   so   P = N E - |Y_0|^2 - |Y_{N/2-1}|^2 - |Y_{N/2-2}|^2:
                // one pass over the row, three DFT bins instead of N.  Lane l holds x[32 a + l]; bin
                // N/2 - m is sum_l (-1)^l e^{-2 pi i l m / N} D_m(l) with D_m(l) bin m of the forward
                // R1B-point DFT over a.  The Doppler window spreads every line over three bins, so
                // some of a row's energy always lies outside the removed bins: >= 27 % on the synthetic
                // sectors (5e-7 relative against the double oracle), 2 % for a line placed exactly between
                // the two clipped bins (3e-5 dB, tests/test_gpu_parity.py).
                const float4 tl = *reinterpret_cast<const float4 *>(tile + Tab::OFF_TL + lane * 16);
                float q[RPW][7];
                static_for<RPW>([&](auto ri) {
                    constexpr int rr = decltype(ri)::value;
                    const uint8_t *row = region + rr * (N * 8);
                    float2 v[R1B];
                    static_for<R1B>([&](auto ai) {
                        constexpr int a = decltype(ai)::value;
                        v[a] = *reinterpret_cast<const float2 *>(row + (32 * a + lane) * 8);
                    });
#ifndef WRP_PV_LATE
                    if constexpr (rr == RPW - 1) { // the warp's region is in registers: fetch the next item
                        publish_next();
                        publish_pending();
                        __syncwarp();
                        prefetch_next(__float_as_uint(v[R1B - 1].x) & (uint32_t)p.zero);
                    }
#endif
                    float2 e2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                    static_for<R1B>([&](auto ai) { // (sum re^2, sum im^2), one FFMA2 per sample
                        constexpr int a = decltype(ai)::value;
                        e2[a & 1] = cfma2(v[a], v[a], e2[a & 1]);
                    });
                    const float2 es = cadd(e2[0], e2[1]);
                    float2 b0, b1, b2;
                    dft_bins012<R1B>(v, b0, b1, b2);
                    const float2 y1 = cmul(b1, make_float2(tl.x, tl.y)), y2 = cmul(b2, make_float2(tl.z, tl.w));
                    q[rr][0] = es.x + es.y;
                    q[rr][1] = b0.x;
                    q[rr][2] = b0.y;
                    q[rr][3] = y1.x;
                    q[rr][4] = y1.y;
                    q[rr][5] = y2.x;
                    q[rr][6] = y2.y;
                });
#ifdef WRP_PV_LATE
                publish_next();
                publish_pending();
                __syncwarp();
                prefetch_next(__float_as_uint(q[RPW - 1][0]) & (uint32_t)p.zero);
#endif
                // lane sums: with two rows per warp the first exchange also transposes, so the low
                // half-warp ends up with row 0 (hh) and the high one with row 1 (vv)
                float r[7];
                if constexpr (RPW == 2) {
                    const bool hi = lane >= 16;
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const float send = hi ? q[0][k] : q[RPW - 1][k], keep = hi ? q[RPW - 1][k] : q[0][k];
                        r[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 7; ++k) r[k] = q[0][k] + __shfl_xor_sync(0xffffffffu, q[0][k], 16);
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
                }
                float removed = r[1] * r[1];
#pragma unroll
                for (int k = 2; k < 7; ++k) removed = fmaf(r[k], r[k], removed);
                pw = fmaxf(fmaf((float)N, r[0], -removed), 0.f);
            } else {
                {
                    const float4 *t4 = reinterpret_cast<const float4 *>(tile + Tab::OFF_TWB + lane * Tab::TWB_ROW);
                    // two rows per warp: the inter-pass twiddles stay in registers across both rows;
                    // one row (N = 1024): they are read as they are used, two steps ahead of the stores
                    float2 tw[RPW == 2 ? R1B : 2];
                    if constexpr (RPW == 2) {
                        static_for<R1B / 2>([&](auto qi) {
                            constexpr int q = decltype(qi)::value;
                            const float4 w = t4[q];
                            tw[2 * q] = make_float2(w.x, w.y);
                            tw[2 * q + 1] = make_float2(w.z, w.w);
                        });
                    }
#pragma unroll 1
                    for (int rr = 0; rr < RPW; ++rr) {
                        uint8_t *row = region + rr * (N * 8);
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
                        if constexpr (RPW == 2) {
                            static_for<R1B>([&](auto ki) {
                                constexpr int ka = decltype(ki)::value;
                                const float2 y = ka == 0 ? v[0] : cmul(v[ka], tw[ka]);
                                *reinterpret_cast<float2 *>(row + (32 * ka + (lane ^ ((ka & 7) << 1))) * 8) = y;
                            });
                        } else {
                            float4 wq[3] = {t4[0], t4[1], t4[2]};
                            static_for<R1B / 2>([&](auto qi) {
                                constexpr int q = decltype(qi)::value;
                                const float4 w = wq[q % 3];
                                if constexpr (q + 3 < R1B / 2) wq[q % 3] = t4[q + 3];
                                const float2 y0 = q == 0 ? v[0] : cmul(v[2 * q], make_float2(w.x, w.y));
                                const float2 y1 = cmul(v[2 * q + 1], make_float2(w.z, w.w));
                                *reinterpret_cast<float2 *>(row + (32 * (2 * q) + (lane ^ (((2 * q) & 7) << 1))) * 8) = y0;
                                *reinterpret_cast<float2 *>(row + (32 * (2 * q + 1) + (lane ^ (((2 * q + 1) & 7) << 1))) * 8) = y1;
                            });
                        }
                    }
                }
                publish_next();
                publish_pending();
                __syncwarp();
                float2 u[32];
                {
                    const uint8_t *grp = region + rsel * (N * 8) + ka * 256;
                    const uint8_t *g_sw[8]; // 16-byte chunk cc of group ka sits at chunk cc ^ (ka & 7)
#pragma unroll
                    for (int k = 0; k < 8; ++k) g_sw[k] = grp + ((k ^ (ka & 7)) * 16);
                    static_for<16>([&](auto ci) {
                        constexpr int cc = decltype(ci)::value;
                        const float4 q = *reinterpret_cast<const float4 *>(g_sw[cc & 7] + (cc & 8) * 16);
                        u[brev<32>(2 * cc)] = make_float2(q.x, q.y);
                        u[brev<32>(2 * cc + 1)] = make_float2(q.z, q.w);
                    });
                }
                __syncwarp();
                dit_stage<32, 1, +1>(u); // consumes every loaded value (see the range tile)
                prefetch_next(__float_as_uint(u[31].x) & (uint32_t)p.zero);
                fft_dit_after_stage1<32, +1>(u);
                // stage 03 shift + clip (rpv2.cu:137-148): the zeroed columns N-1, N-2 are bins N/2-1 =
                // (R1B-1) + R1B*15 and N/2-2; stage 04 |.|^2 and the row sum (rpv2.cu:150-157, 171-197)
                if (ka >= R1B - 2) u[15] = make_float2(0.f, 0.f); // the two clipped bins
                float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                static_for<32>([&](auto ki) { // (sum re^2, sum im^2) with one FFMA2 per bin
                    constexpr int kb = decltype(ki)::value;
                    acc[kb & 3] = cfma2(u[kb], u[kb], acc[kb & 3]);
                });
                const float2 a2 = cadd(cadd(acc[0], acc[1]), cadd(acc[2], acc[3]));
                pw = a2.x + a2.y;
#pragma unroll
                for (int o = R1B / 2; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
            }
            pw *= p.taps_sum; // stages 05-08: row sum of the circular convolution
            // (channel, gate) of this thread's row — the same map the loads use
            int chn, gate;
            (void)doppler_row<N, NW>(it, p, warp, rsel, chn, gate);
            if (p.power && ka == 0) p.power[((size_t)it.sector * p.C + chn) * p.half_m + gate] = pw;
            if constexpr (RPW == 2) {
                // stages 09/10 (rpv2.cu:199-213): hh in the low half-warp, vv in the high one
                const float other = __shfl_xor_sync(0xffffffffu, pw, 16);
                if ((pair && lane == 0) || (!pair && p.C == 1 && ka == 0)) {
                    const float rg = (float)gate * p.range_res;
                    const float z = rg * rg * p.calib * pw;
                    reinterpret_cast<float2 *>(p.out)[(size_t)it.sector * p.half_m + gate] =
                        make_float2(10.f * log10f(z), pair ? 10.f * (log10f(pw) - log10f(other)) : 0.f);
                }
            } else {
                // N = 1024: one row per warp, (hh, vv) of a gate sit in neighbouring warps
                if (lane == 0) p_row[warp] = pw;
                __syncthreads();
                if (lane == 0 && ((pair && !(warp & 1)) || (!pair && p.C == 1))) {
                    const float rg = (float)gate * p.range_res;
                    const float z = rg * rg * p.calib * pw;
                    reinterpret_cast<float2 *>(p.out)[(size_t)it.sector * p.half_m + gate] =
                        make_float2(10.f * log10f(z), pair ? 10.f * (log10f(pw) - log10f(p_row[warp + 1])) : 0.f);
                }
                __syncthreads();
            }
        }

        if (nit.kind < 0) {
            publish_pending(); // leaving: nothing may stay unpublished
        } else if (!loaded) {
            // the next item's dependency was unmet when probed.  Publish what this warp still holds
            // (the item waited for may be this CTA's own), then thread 0 waits for the counter and
            // releases everybody through shared memory.
            publish_pending();
            if (tid == 0) {
                int target;
                const int *dep = item_dep<T>(nit, p, target);
                if (dep) spin_until(dep, target);
                s_go = n + 2;
            }
            while (s_go < n + 2) {
            }
            issue_warp_load<N, T, Q>(nit, p, tile, &mbar, warp, lane);
        }
        it = nit;
        ++n;
        claimed_next = claimed_next2; // consumed at the next item's top: the atomic has long returned
    }
}

// ---- host side ---------------------------------------------------------------------------------
bool persistent_supported(int M, int N) { return (M == 1024 || M == 4096) && (N == 512 || N == 1024); }
int persistent_ctrl_ints(int smax) { return CTRL_A + 2 * smax; }

static int tile_cols(int M)
{
    // 8 columns (64-byte row segments) measured 24 % faster than 4 for M = 1024.
    // 4096 rows: 4 columns = 128 KiB tiles, one 16-warp CTA per SM.  (2 columns = 64 KiB tiles with two
    // 8-warp CTAs per SM was measured 20 % slower: 16-byte row pieces, 10x the bank conflicts.)
    return M == 4096 ? 4 : 8;
}

cudaError_t persistent_setup()
{
    cudaError_t e;
#define WRP_SET(R1B, T, Q)                                                                                   \
    e = cudaFuncSetAttribute(chain_persistent_kernel<R1B, T, Q, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             Tables<R1B, T, Q>::SMEM);                                                       \
    if (e != cudaSuccess) return e;                                                                          \
    e = cudaFuncSetAttribute(chain_persistent_kernel<R1B, T, Q, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             Tables<R1B, T, Q>::SMEM);                                                       \
    if (e != cudaSuccess) return e;
    WRP_SET(16, 8, 1)
    WRP_SET(32, 8, 1)
    WRP_SET(16, 4, 4)
    WRP_SET(32, 4, 4)
#undef WRP_SET
    return cudaSuccess;
}

// One launch for the whole batch.  ctrl must hold CTRL_A + 2*smax ints.  The atomic work queue hands
// items out in dependency order and a CTA only ever waits for items claimed earlier, so the kernel
// terminates for any number of resident CTAs (the grid is sized from the occupancy query anyway).
cudaError_t launch_persistent(const float2 *iq, float *out, float *power, float2 *x2_ring, int ring, int lag,
                              int *ctrl, int smax, const FusedTables &t, int M, int N, int C, int n_sectors,
                              float range_res, float calib, float taps_sum, int sm_count, bool doppler_fft,
                              int evict_first, int debug, cudaStream_t st)
{
    if (n_sectors == 0) return cudaSuccess;
    if (!persistent_supported(M, N) || n_sectors > smax || ring < lag + 2) return cudaErrorInvalidValue;
    const int T = tile_cols(M), Q = M / 1024, NW = T * Q;

    PersistParams p{};
    p.wrc_t = t.wrc_t;
    p.wd = t.wd;
    p.tw_a = t.tw_a;
    p.tw_b = t.tw_b;
    p.wr4 = t.wr4;
    p.tw4 = t.tw4;
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
    const int rows_b = NW * 1024 / N;
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
    // (M = 4096: the hand-off of even one sector exceeds L2, so protecting it buys nothing; measured 5 % slower)
    p.evict_first = evict_first >= 0 ? evict_first : (M == 1024);
    p.discard = 0;
    p.debug = debug;
    p.range_res = range_res;
    p.calib = calib;
    p.taps_sum = taps_sum;

    cudaError_t e = cudaMemsetAsync(ctrl, 0, sizeof(int) * (CTRL_A + 2 * (size_t)smax), st);
    if (e != cudaSuccess) return e;
    int grid = (16 / NW) * sm_count;
    if (grid > p.total_items) grid = p.total_items;
#define WRP_LAUNCH(R1B, TT, QQ)                                                                              \
    do {                                                                                                     \
        if (doppler_fft)                                                                                     \
            chain_persistent_kernel<R1B, TT, QQ, 0><<<grid, 32 * NW, Tables<R1B, TT, QQ>::SMEM, st>>>(p);   \
        else                                                                                                 \
            chain_persistent_kernel<R1B, TT, QQ, 1><<<grid, 32 * NW, Tables<R1B, TT, QQ>::SMEM, st>>>(p);   \
        return cudaGetLastError();                                                                           \
    } while (0)
    if (Q == 4) {
        if (N == 512) WRP_LAUNCH(16, 4, 4);
        WRP_LAUNCH(32, 4, 4);
    }
    if (N == 512) WRP_LAUNCH(16, 8, 1);
    WRP_LAUNCH(32, 8, 1);
#undef WRP_LAUNCH
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
