// wrp_stream.cu — the fused chain as ONE persistent kernel with NO range -> Doppler hand-off (sm_100a).
//
// Stages 03-08 in energy form need, per range gate k of a (sector, channel) plane, only four sums over
// the Doppler axis j of the range-FFT output x2[k][j]  (rpv2.cu:93-197 collapsed by Parseval, see
// DESIGN.md section 4.0):
//     E   = sum_j |x2[k][j]|^2
//     Y_0 = sum_j x2[k][j]                                   (the bin the mean removal zeroes)
//     Y_m = sum_j (-1)^j exp(-2 pi i j m / N) x2[k][j], m = 1, 2   (the two clipped bins N/2 - m)
//     P[k] = (N E - |Y_0|^2 - |Y_1|^2 - |Y_2|^2) * sum(taps)
// A range tile (T adjacent columns x M rows) contributes T terms to each sum of each of its M/2
// surviving rows.  So the tile's epilogue folds its rows straight into seven accumulators per gate and
// the range-FFT output never leaves the SM: no x2 ring in L2, no second kernel phase, no dependency
// between CTAs, DRAM traffic = the input, L2 traffic = the input.
//
// One CTA walks a contiguous run of tiles of one plane after another, carrying the accumulators of
// the current plane (registers for M = 1024, shared memory for M = 4096):
//   load   the tile, a whole tile ahead: planar input by TMA (cp.async.bulk.tensor, one 8 KiB box per
//          warp-local region of the tile buffer, issued the moment the warp has pulled its last operand
//          out of it); wire input: 4-byte cp.async of the channel's int16 pair out of every 12-byte record into a
//          separate landing buffer, big-endian decode in the first pass — sector.cpp:52-62 on the load path)
//   pass 1 window folded into the first butterfly stage (rpv2.cu:86-91), radix-32 in registers,
//          inter-pass twiddle, in-place exchange through shared memory          (stages 01-02,
//   pass 2 radix-32, outputs k < M/2 only                                        rpv2.cu:318-333,426-428)
//   fold   outputs staged warp-locally as [gate][T columns], each lane reads whole rows back and
//          updates E, Y_0, Y_1, Y_2 of its gates (no shuffles, no CTA barrier)
//   plane end: P[k] -> power[]; the CTA that completes the second of (hh, vv) writes ZdB/ZDR
//          (rpv2.cu:199-213).  A plane cut by the work partition is summed by the last CTA to arrive
//          (partials in scratch, fixed order: results do not depend on arrival order).
// Nothing in this kernel waits for another CTA: any grid size is correct, co-residency is irrelevant.
#include "wrp_stream.h"
#include <cstdlib>

#include "wrp_fft.cuh"
#include "wrp_internal.h"
#include "wrp_ptx.cuh"

namespace wrp {
namespace stream {

// Checked build (make checked -> tools/libwrp_checked.so, -DWRP_CHECKED): every index a kernel of this file uses
// for a global store, and every tile / plane / sector it derives from its work partition, is range-checked and a
// violation traps (-> cudaErrorLaunchFailure -> WRP_ERR_CUDA).  compute-sanitizer is closed on the measurement pool,
// so this is how the GPU tests double as a memory-safety run (profiles/r02_checked_build.txt).  No code in a
// normal build.
#ifdef WRP_CHECKED
#define WRP_CHECK(cond)                                                                                              \
    do {                                                                                                             \
        if (!(cond)) __trap();                                                                                       \
    } while (0)
#else
#define WRP_CHECK(cond) ((void)0)
#endif

constexpr int R = 32;
constexpr int WRC_ROW = 32 * 4 + 16; // wr(i)*c transposed [32 b][32 a] floats, rows padded by 16 B
constexpr int TWA_ROW = 32 * 8 + 16; // range inter-pass twiddles [32 b][32 ka] float2

template <int Q, bool WIRE> struct Cfg {
    static_assert(Q == 1 || (Q == 4 && !WIRE), "M = 1024 (planar or wire) or M = 4096 (planar)");
    static constexpr int T = Q == 1 ? 8 : 4;      // columns per tile
    static constexpr int NW = T * Q;              // warps: one per 8 KiB of the exchange buffer
    static constexpr int THREADS = 32 * NW;
    static constexpr int PITCH = T * 8;           // bytes per row of the exchange buffer
    static constexpr int CPR = PITCH / 16;        // 16-byte chunks per row
    static constexpr int XBUF = 1024 * Q * PITCH; // exchange buffer; also the landing zone of planar input
    static constexpr int LAND = WIRE ? 1024 * T * 4 : 0; // wire: (I, Q) int16 pairs of the tile
    static constexpr int KPW = 32 / T;            // ka values per warp
    // the fold runs in PHASES rounds of 16 / PHASES output rows kb: a warp stages 32 gates per round
    // (KB_PHASE kb x KPW ka) and every lane reads one of them back
    static constexpr int PHASES = Q == 1 ? 2 : 4;
    static constexpr int KB_PHASE = 16 / PHASES;
    static constexpr int ROWS_PHASE = KB_PHASE * KPW;
    static_assert(ROWS_PHASE == 32, "one staged gate per lane and round");
    // planar: a separate staging buffer, because the warp's region of the exchange buffer is already
    // receiving the next tile; wire: the region itself (the next tile lands in the landing buffer)
    static constexpr bool STAGE_SEPARATE = !WIRE;
    static constexpr int STAGE = STAGE_SEPARATE ? NW * ROWS_PHASE * PITCH : 0;
    static constexpr int RPT = PHASES;            // gates per thread: 2 (M = 1024) or 4 (M = 4096)
    static constexpr bool ACC_SMEM = Q == 4;
    static constexpr int ACC = ACC_SMEM ? RPT * 2 * THREADS * 16 : 0; // [gate slot][chunk][thread] float4
    static constexpr int OFF_LAND = XBUF;
    static constexpr int OFF_STAGE = OFF_LAND + LAND;
    static constexpr int OFF_ACC = OFF_STAGE + STAGE;
    static constexpr int OFF_TAB = OFF_ACC + ACC;
    static constexpr int OFF_WRC = OFF_TAB;                      // Q = 1
    static constexpr int OFF_TW4 = OFF_TAB;                      // Q = 4: exp(-2 pi i r / 4096) [1024]
    // (Q = 4: the window wr(i)*c [4096] stays in global memory — a thread needs the same 16 entries for every
    // tile and fetches them before it waits for the tile; shared memory holds tile + sums + staging)
    static constexpr int OFF_TWA = Q == 1 ? OFF_WRC + 32 * WRC_ROW : OFF_TW4 + 1024 * 8;
    static constexpr int SMEM = OFF_TWA + 32 * TWA_ROW;
    static_assert((Q == 1 ? 2 : 1) * (SMEM + 1024 + 64) <= 233472, "shared memory per SM");
};

// Stage 09/10 store (rpv2.cu:199-213 writes result[] of the sector; rpv2.cu:607, 736 collect the volume).  With
// product mirrors set the kernel is its own gather: the (ZdB, ZDR) pair also goes to the same index of every
// mirror — peer-mapped buffers of the other devices over NVLink — 4 KiB per sector and peer against 12.6 MB read,
// fire-and-forget stores from the epilogue, no collective kernel and no copy afterwards.
__device__ __forceinline__ void store_product(const StreamParams &p, size_t idx, float2 v)
{
    reinterpret_cast<float2 *>(p.out)[idx] = v;
    for (int m = 0; m < p.n_mirrors; ++m) reinterpret_cast<float2 *>(p.mirror[m])[idx] = v;
}

// One staged row of T columns x[0..T) (already scaled by wd(j)) folded into a gate's seven sums.
//   E += sum |x_c|^2,  Y_0 += sum x_c,  Y_m += tw_m * S_m  with  S_m = sum_c (-1)^c e^{-i theta_m c} x_c.
// Columns are paired around the tile's centre: with P = x_c + x_{T-1-c}, D = x_c - x_{T-1-c},
//   (-1)^c e^{-i theta c} x_c + (-1)^{T-1-c} e^{-i theta (T-1-c)} x_{T-1-c}
//       = e^{-i theta (T-1)/2} (-1)^c [cos(phi) D + i sin(phi) P],   phi = theta ((T-1)/2 - c),
// so a pair costs four real FMAs per bin instead of eight, and the common factor lives in tile_tw.
// kp[m][c] = ((-1)^c cos(phi_{m,c}), (-1)^c sin(phi_{m,c})) comes from the kernel's constant bank.
template <int T>
__device__ __forceinline__ void fold_row(const float2 (&x)[T], const float2 (&kp)[2][8], const float4 ttw, float (&a7)[7])
{
    float2 e2 = cmul2(x[0], x[0]);
    static_for<T - 1>([&](auto ci) {
        constexpr int cc = decltype(ci)::value + 1;
        e2 = cfma2(x[cc], x[cc], e2);
    });
    float2 P[T / 2], D[T / 2];
    static_for<T / 2>([&](auto pi) {
        constexpr int q = decltype(pi)::value;
        P[q] = cadd(x[q], x[T - 1 - q]);
        D[q] = csub(x[q], x[T - 1 - q]);
    });
    float2 s0 = P[0];
    static_for<T / 2 - 1>([&](auto pi) {
        constexpr int q = decltype(pi)::value + 1;
        s0 = cadd(s0, P[q]);
    });
    float2 s[2];
    static_for<2>([&](auto mi) {
        constexpr int m = decltype(mi)::value;
        float sx = kp[m][0].x * D[0].x, sy = kp[m][0].x * D[0].y;
        sx = fmaf(-kp[m][0].y, P[0].y, sx);
        sy = fmaf(kp[m][0].y, P[0].x, sy);
        static_for<T / 2 - 1>([&](auto pi) {
            constexpr int q = decltype(pi)::value + 1;
            sx = fmaf(kp[m][q].x, D[q].x, sx);
            sx = fmaf(-kp[m][q].y, P[q].y, sx);
            sy = fmaf(kp[m][q].x, D[q].y, sy);
            sy = fmaf(kp[m][q].y, P[q].x, sy);
        });
        s[m] = make_float2(sx, sy);
    });
    a7[0] += e2.x + e2.y;
    a7[1] += s0.x;
    a7[2] += s0.y;
    a7[3] = fmaf(s[0].x, ttw.x, a7[3]);
    a7[3] = fmaf(-s[0].y, ttw.y, a7[3]);
    a7[4] = fmaf(s[0].x, ttw.y, a7[4]);
    a7[4] = fmaf(s[0].y, ttw.x, a7[4]);
    a7[5] = fmaf(s[1].x, ttw.z, a7[5]);
    a7[5] = fmaf(-s[1].y, ttw.w, a7[5]);
    a7[6] = fmaf(s[1].x, ttw.w, a7[6]);
    a7[6] = fmaf(s[1].y, ttw.z, a7[6]);
}

// ---- tile loads ------------------------------------------------------------------------------
// planar: TMA.  The warp's own 8 KiB region of the tile (rows [8192/PITCH * warp, ...)) is one box
// {T columns x 8192/PITCH rows} of the batch viewed as a [planes * M][N] matrix of 8-byte elements:
// lane 0 posts the byte count on the tile barrier and issues one cp.async.bulk.tensor — no LDGSTS, no
// per-lane address arithmetic, and the shared-memory write side does not go through the LSU
// (tools/micro/tile_load.cu: 64-byte-row boxes sustain the same 6.8 TB/s as cp.async).
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// `after_and_zero`: (a register produced by the LAST shared-memory load the issuing thread made from the destination
// buffer) AND (StreamParams::zero, a kernel parameter that is always 0 but that ptxas cannot know).  Added to the row
// coordinate it changes nothing and makes the copy data-dependent on that load in the SASS, so it cannot issue before
// the loads of the old tile have returned (loads of one warp return in order).  Without it the copy is only ordered
// behind the ISSUE of those loads — and a box that hits L2 can overtake loads queued in a busy LSU and overwrite the
// rows they are about to read (seen as run-to-run differences in the first rows of a warp's region of the tile).
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                            uint32_t after_and_zero = 0)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1 + (int)after_and_zero)
                 : "memory");
}
template <int Q, bool WIRE>
__device__ __forceinline__ void issue_tile_planar(const CUtensorMap *tmap, uint8_t *xbuf, uint64_t *bar, int plane, int t,
                                                  int warp, int lane, uint32_t after)
{
    using K = Cfg<Q, WIRE>;
    constexpr int RPWARP = 8192 / K::PITCH;
    if (lane == 0) {
        mbar_expect_tx(bar, 8192);
        tma_load_2d(xbuf + warp * 8192, tmap, t * K::T, plane * (1024 * Q) + warp * RPWARP, bar, after);
    }
}
// wire: every thread fetches 32 of the tile's 8192 (I, Q) pairs — 4 bytes at offset 4 ch of the
// 12-byte record (sector.cpp:52-62) — into the dense landing buffer [row][8 columns]
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void issue_tile_wire(const StreamParams &p, uint8_t *land, uint64_t *bar, int sector, int ch,
                                                int t, int tid)
{
    const size_t row_bytes = (size_t)p.N * 12;
    const int row = tid >> 3, c = tid & 7;
    const uint8_t *src =
        (const uint8_t *)p.in + ((size_t)sector * 1024 + row) * row_bytes + ((size_t)t * 8 + c) * 12 + ch * 4;
    uint8_t *dst = land + tid * 4;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) cp_async4(dst + k * 1024, src + (size_t)k * 32 * row_bytes);
    cp_async_arrive(bar);
}

// ---- the kernel ------------------------------------------------------------------------------
template <int Q, bool WIRE>
__global__ void __launch_bounds__(Cfg<Q, WIRE>::THREADS, Q == 1 ? 2 : 1)
    chain_stream_kernel(const StreamParams p, const __grid_constant__ CUtensorMap tmap)
{
    using K = Cfg<Q, WIRE>;
    constexpr int T = K::T, NW = K::NW, THREADS = K::THREADS, PITCH = K::PITCH, CPR = K::CPR, KPW = K::KPW;
    constexpr int RPT = K::RPT;
    constexpr int SW = 128 / PITCH - 1; // row-swizzle mask of the in-place exchange
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar; // the next tile has landed: planar — one arrival + 8 KiB of TMA bytes per warp;
                                           // wire — one arrival per thread, fired by its cp.asyncs
    __shared__ __align__(8) uint64_t ebar; // wire: every warp is done with its staged rows of the previous tile
    __shared__ int s_flag;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *const xbuf = smem;

    // ---- work partition: a contiguous run of tiles of the [virtual plane][tile] space --------------
    const int CG = p.chan_groups;
    const int grp = blockIdx.x % CG, u = blockIdx.x / CG, Gp = gridDim.x / CG;
    if (u >= Gp) return;
    const long long total = (long long)(CG == 1 ? p.S * p.C : p.S) * p.NT;
    const int g_lo = (int)((long long)u * total / Gp), g_end = (int)(((long long)u + 1) * total / Gp);
    if (g_lo >= g_end) return;
    auto real_plane = [&](int vp) { return CG == 1 ? vp : vp * p.C + grp; };
    auto cta_of = [&](long long g) { return (int)(((g + 1) * Gp - 1) / total); }; // u of the CTA that owns tile g

    // ---- tables -> shared memory -----------------------------------------------------------------
    {
        auto copy_rows = [&](int off, const void *src, int rows, int row_bytes, int row_pitch) {
            const int per_row = row_bytes / 16;
            for (int i = tid; i < rows * per_row; i += THREADS) {
                const int r = i / per_row, q = i - r * per_row;
                *reinterpret_cast<float4 *>(smem + off + r * row_pitch + q * 16) =
                    __ldg(reinterpret_cast<const float4 *>(src) + i);
            }
        };
        if constexpr (Q == 1) {
            copy_rows(K::OFF_WRC, p.wrc_t, 32, 32 * 4, WRC_ROW);
        } else {
            copy_rows(K::OFF_TW4, p.tw4, 1, 1024 * 8, 1024 * 8);
        }
        copy_rows(K::OFF_TWA, p.tw_a, 32, 32 * 8, TWA_ROW);
        if constexpr (K::ACC_SMEM) {
            for (int i = tid; i < K::ACC / 16; i += THREADS)
                *reinterpret_cast<float4 *>(smem + K::OFF_ACC + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (tid == 0) {
        mbar_init(&mbar, WIRE ? THREADS : NW);
        mbar_init(&ebar, NW);
    }
    __syncthreads();

    int vp = g_lo / p.NT, t = g_lo - vp * p.NT; // virtual plane and tile of the current item
    const int vp_lo = vp;
    auto issue_tile = [&](int vplane, int tile, uint32_t after = 0) {
        const int plane = real_plane(vplane);
        if constexpr (WIRE) {
            const int sector = plane / p.C;
            issue_tile_wire(p, smem + K::OFF_LAND, &mbar, sector, plane - sector * p.C, tile, tid);
        } else {
            issue_tile_planar<Q, WIRE>(&tmap, xbuf, &mbar, plane, tile, warp, lane, after);
        }
    };
    issue_tile(vp, t);

    // ---- per-thread constants of the fold ---------------------------------------------------------
    // writer: thread (column c, ka_l) stores output kb of the phase at staged row R = kbl * KPW + ka_l;
    // 16-byte chunk (c >> 1) of a row is XOR-swizzled with s(R) so that the 128-bit row reads below are
    // conflict-free: s(R) = (R >> 1) & 3 for 64-byte rows, (R >> 2) & 1 for 32-byte rows
    const int cw = lane % T, ka_l = lane / T;
    uint8_t *const stage = K::STAGE_SEPARATE ? smem + K::OFF_STAGE + warp * (K::ROWS_PHASE * PITCH) : xbuf + warp * 8192;
    uint8_t *wbase[2]; // by parity of kbl (T = 8: s depends on it; T = 4: both entries equal)
    static_assert(T == 8 || (KPW * 4) % 8 == 0, "T = 4: s(R) = (R >> 2) & 1 must not depend on kbl");
#pragma unroll
    for (int par = 0; par < 2; ++par) {
        const int s = T == 8 ? ((par << 1) | (ka_l >> 1)) : ((ka_l >> 2) & 1);
        wbase[par] = stage + ka_l * PITCH + (((cw >> 1) ^ s) << 4) + (cw & 1) * 8;
    }
    // reader: lane reads staged row R = lane of the round; logical chunk q sits at q ^ s(R)
    const int rs = T == 8 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);
    const uint8_t *const rbase = stage + lane * PITCH;
    // gate of fold slot r (= round) of this thread: staged row lane = kbl * KPW + ka_l
    auto slot_gate = [&](int r) {
        if constexpr (Q == 1) {
            const int kb = 8 * r + (lane >> 2), ka = 4 * warp + (lane & 3);
            return ka + 32 * kb;
        } else {
            const int kb = 4 * r + (lane >> 3), ka = 8 * (warp & 3) + (lane & 7);
            return 4 * (ka + 32 * kb) + (warp >> 2);
        }
    };
    float acc[K::ACC_SMEM ? 1 : RPT][7]; // E, Re/Im Y_0, Re/Im Y_1, Re/Im Y_2 (registers, M = 1024)
    if constexpr (!K::ACC_SMEM) {
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int q = 0; q < 7; ++q) acc[r][q] = 0.f;
    }
    float4 *const acc_s = reinterpret_cast<float4 *>(smem + K::OFF_ACC) + tid; // [slot][chunk][thread]

    uint32_t phase = 0, ephase = 0;
    bool first = true;

    for (int g = g_lo; g < g_end; ++g) {
        int nt = t + 1, nvp = vp;
        if (nt == p.NT) nt = 0, ++nvp;
        const bool has_next = g + 1 < g_end;
        const int plane = real_plane(vp);
        WRP_CHECK(t >= 0 && t < p.NT && plane >= 0 && plane < p.S * p.C && g >= 0 && g < total);
        const float4 ttw = __ldg(p.tile_tw + t); // this tile's factors of the two clipped bins

        // thread -> (sub-tile, column, b): Q = 4: thread group `sub` (32 T threads) owns the 1024-row sub-tile k0 = sub
        const int sub = Q == 1 ? 0 : tid / (32 * T), tl = Q == 1 ? tid : tid % (32 * T);
        const int c = tl % T, b = tl / T;
        const int col0 = t * T, col = col0 + c;
        uint8_t *const stile = xbuf + sub * (1024 * PITCH);
        float wdj = 0.f;
        if constexpr (Q == 1) wdj = __ldg(p.wd + col);
        float2 wdp = make_float2(0.f, 0.f);
        if constexpr (Q == 4) wdp = __ldg(reinterpret_cast<const float2 *>(p.wd + col0) + tid % (T / 2));

        // M = 4096: the 16 window values of this thread's pre-pass units (the same for every tile), fetched
        // from global memory while the tile is still in flight
        float wr4v[Q == 4 ? 4 : 1][4];
        if constexpr (Q == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) wr4v[k][q] = __ldg(p.wr4 + (tid + k * THREADS) / (T / 2) + 1024 * q);
        }

        mbar_wait(&mbar, phase);
        phase ^= 1;

        if constexpr (Q == 4) {
            // stage 01 + radix-4 DIF step over rows r, r + 1024, r + 2048, r + 3072, in place:
            //   y_k0[r] = W_4096^(r k0) * sum_q (-i)^(q k0) ham(r + 1024 q) x[r + 1024 q]
            // one unit = one row r x two adjacent columns (16-byte accesses); the 1024-point transforms of
            // the four sub-tiles then yield rows 4 k' + k0 of the 4096-point transform
            constexpr int UNITS = 1024 * T / 2;
            static_assert(UNITS == 4 * THREADS, "four pre-pass units per thread");
            const int cp = tid % (T / 2);
            const float2 *tw4 = reinterpret_cast<const float2 *>(smem + K::OFF_TW4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int un = tid + k * THREADS;
                const int r = un / (T / 2);
                uint8_t *ptr = xbuf + r * PITCH + cp * 16;
                float4 x[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) x[q] = *reinterpret_cast<const float4 *>(ptr + q * (1024 * PITCH));
                float2 e[4], o[4]; // even / odd column of the pair
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float w = wr4v[k][q];
                    const float wa = w * wdp.x, wb = w * wdp.y;
                    e[q] = cmul2(make_float2(x[q].x, x[q].y), make_float2(wa, wa));
                    o[q] = cmul2(make_float2(x[q].z, x[q].w), make_float2(wb, wb));
                }
                const float2 w1 = tw4[r], w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                auto radix4 = [&](float2(&z)[4]) {
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
        }

        // ================= pass 1: rows 32 a + b of column c =================
        float2 v[R];
        if constexpr (WIRE) {
            // big-endian int16 (I, Q) -> float: one PRMT per component swaps the bytes and extends the sign
            const uint8_t *src = smem + K::OFF_LAND + b * (T * 4) + c * 4;
            static_for<R>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                const uint32_t w = *reinterpret_cast<const uint32_t *>(src + a * (R * T * 4));
                // memory bytes I_hi I_lo Q_hi Q_lo = bytes 0..3 of w: I = sext(b0 : b1), Q = sext(b2 : b3)
                v[brev<R>(a)] = make_float2((float)(int)prmt(w, 0u, 0x8801u), (float)(int)prmt(w, 0u, 0xAA23u));
            });
        } else {
            const uint8_t *src = stile + b * PITCH + c * 8;
            static_for<R>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                v[brev<R>(a)] = *reinterpret_cast<const float2 *>(src + a * (R * PITCH));
            });
        }
        if constexpr (Q == 1) {
            // stage 01 (x *= wr(i)*c*wd(j), rpv2.cu:86-91) fused into the first butterfly stage: the span-1
            // partners of the bit-reversed network are rows a and a + 16
            const float2 m2 = make_float2(-2.f, -2.f);
            const float4 *w4 = reinterpret_cast<const float4 *>(smem + K::OFF_WRC + b * WRC_ROW);
            static_for<R / 8>([&](auto qi) { // rows 4q .. 4q+3 and their partners 16 + 4q ..
                constexpr int q = decltype(qi)::value;
                const float4 wlo = w4[q], whi = w4[q + R / 8];
                const float lo[4] = {wlo.x, wlo.y, wlo.z, wlo.w}, hi[4] = {whi.x, whi.y, whi.z, whi.w};
                static_for<4>([&](auto ei) {
                    constexpr int e = decltype(ei)::value;
                    constexpr int sa = brev<R>(4 * q + e); // even slot; partner row a + R/2 sits in sa + 1
                    static_assert(brev<R>(4 * q + e + R / 2) == sa + 1, "span-1 partner");
                    // (wd(j) applied once to the 16 pass-2 outputs instead — 16 packed products for 32 scalar ones —
                    // was measured 1.5 % SLOWER: it lengthens the chain between the transform and the fold)
                    const float wl = lo[e] * wdj, wh = hi[e] * wdj;
                    const float2 tt = cmul2(v[sa + 1], make_float2(wh, wh));
                    const float2 s2 = cfma2(v[sa], make_float2(wl, wl), tt); // A*wl + B*wh
                    v[sa + 1] = cfma2(tt, m2, s2);                            // A*wl - B*wh
                    v[sa] = s2;
                });
            });
            fft_dit_after_stage1<R, -1>(v);
        } else {
            fft_dit<R, -1>(v);
        }
        if constexpr (WIRE) {
            // the exchange buffer doubles as the fold's staging area: wait until every warp has read its
            // staged rows of the previous tile back before overwriting them
            if (!first) {
                mbar_wait(&ebar, ephase);
                ephase ^= 1;
            }
        }
        {
            // Z[ka][b] goes to row 32 ka + (b ^ (ka & SW)): the warp keeps its own row footprint (in place);
            // the twiddle reads run two steps ahead of the exchange stores
            const float4 *t4 = reinterpret_cast<const float4 *>(smem + K::OFF_TWA + b * TWA_ROW);
            uint8_t *d_sw[SW + 1];
#pragma unroll
            for (int sx = 0; sx <= SW; ++sx) d_sw[sx] = stile + (b ^ sx) * PITCH + c * 8;
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
        // the exchange: the one barrier of a 1024-point column group
        if constexpr (Q == 1) {
            __syncthreads();
        } else { // constant barrier ids, so that only 5 of the 16 are reserved
            if (sub == 0) bar_sync_id<2>(32 * T);
            else if (sub == 1) bar_sync_id<3>(32 * T);
            else if (sub == 2) bar_sync_id<4>(32 * T);
            else bar_sync_id<5>(32 * T);
        }
        if constexpr (WIRE) {
            if (has_next) issue_tile(nvp, nt); // the landing buffer is free: every thread is past its first pass
        }

        // ================= pass 2: rows 32 ka + b' of column c, ka = b =================
        const int ka = b; // rows 32 ka + .. of warp w are its own 8 KiB region
        {
            const uint8_t *s_sw[SW + 1];
#pragma unroll
            for (int sx = 0; sx <= SW; ++sx) s_sw[sx] = stile + ka * (R * PITCH) + c * 8 + ((ka ^ sx) & SW) * PITCH;
            static_for<R>([&](auto bi) {
                constexpr int bb = decltype(bi)::value;
                v[brev<R>(bb)] = *reinterpret_cast<const float2 *>(s_sw[bb & SW] + (bb & ~SW) * PITCH);
            });
        }
        __syncwarp();
        // first butterfly stage: it consumes every loaded value, so what follows is ordered behind the COMPLETION of
        // this warp's pass-2 reads (tma_load_2d / mbar_arrive_after explain why their issue is not enough)
        dit_stage<R, 1, -1>(v);
        {
            const uint32_t dep = __float_as_uint(v[R - 1].x) & (uint32_t)p.zero;
            if constexpr (K::STAGE_SEPARATE) {
                // the warp's region is in registers: fetch its share of the next tile.  (Issuing after the second /
                // third / last butterfly stage instead was measured 2-3 % slower; a TMA L2 prefetch of the tile
                // after next +0.5 % on 1024 x 512 and -12 % on 4096 x 1024.)
                if (has_next) issue_tile(nvp, nt, dep);
            }
        }
        fft_dit_after_stage1<R, -1>(v);
        if (p.x2_tap) { // debug tap (tests): stage 02 rows k < M/2 as the product kernel computes them
            WRP_CHECK(plane >= 0 && plane < p.S * p.C && Q * ka + sub + Q * R * 15 < p.half_m && col >= 0 && col < p.N);
            float2 *o = p.x2_tap + ((size_t)plane * p.half_m + Q * ka + sub) * (size_t)p.N + col;
            static_for<R / 2>([&](auto ki) {
                constexpr int kb = decltype(ki)::value;
                o[(size_t)(Q * R * kb) * p.N] = v[kb];
            });
        }

        // ================= fold: stages 03-08 in energy form =================
        // output kb of (c, ka) is gate Q (ka + 32 kb) + sub, column col.  PHASES rounds: stage 32 of the warp's
        // output rows as [gate][T columns], every lane reads one whole row back and updates the gate's seven sums.
        static_for<K::PHASES>([&](auto hi_) {
            constexpr int h = decltype(hi_)::value;
            static_for<K::KB_PHASE>([&](auto ki) {
                constexpr int kbl = decltype(ki)::value;
                *reinterpret_cast<float2 *>(wbase[kbl & 1] + kbl * (KPW * PITCH)) = v[K::KB_PHASE * h + kbl];
            });
            __syncwarp();
            {
                constexpr int slot = h;
                float2 x[T];
                static_for<CPR>([&](auto qi) {
                    constexpr int q = decltype(qi)::value;
                    const float4 w = *reinterpret_cast<const float4 *>(rbase + ((q ^ rs) << 4));
                    x[2 * q] = make_float2(w.x, w.y);
                    x[2 * q + 1] = make_float2(w.z, w.w);
                });
                float a7[7];
                if constexpr (K::ACC_SMEM) {
                    const float4 lo = acc_s[(slot * 2 + 0) * THREADS], hi4 = acc_s[(slot * 2 + 1) * THREADS];
                    a7[0] = lo.x, a7[1] = lo.y, a7[2] = lo.z, a7[3] = lo.w, a7[4] = hi4.x, a7[5] = hi4.y, a7[6] = hi4.z;
                } else {
#pragma unroll
                    for (int q = 0; q < 7; ++q) a7[q] = acc[slot][q];
                }
                fold_row<T>(x, p.wcol, ttw, a7);
                if constexpr (K::ACC_SMEM) {
                    acc_s[(slot * 2 + 0) * THREADS] = make_float4(a7[0], a7[1], a7[2], a7[3]);
                    acc_s[(slot * 2 + 1) * THREADS] = make_float4(a7[4], a7[5], a7[6], 0.f);
                } else {
#pragma unroll
                    for (int q = 0; q < 7; ++q) acc[slot][q] = a7[q];
                }
            }
            __syncwarp();
        });
        if constexpr (WIRE) { // behind the completion of the last read-back (its sums), not its issue
            if (lane == 0) mbar_arrive_after(&ebar, __float_as_uint(acc[K::ACC_SMEM ? 0 : K::PHASES - 1][0]) & (uint32_t)p.zero);
        }

        // ================= end of the plane (or of this CTA's run): products =================
        if (nt == 0 || !has_next) {
            const int hm = p.half_m;
            const long long plane_first = (long long)vp * p.NT;
            const bool complete = g_lo <= plane_first && nt == 0;
            float vals[RPT][7];
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                if constexpr (K::ACC_SMEM) {
                    const float4 lo = acc_s[(r * 2 + 0) * THREADS], hi4 = acc_s[(r * 2 + 1) * THREADS];
                    vals[r][0] = lo.x, vals[r][1] = lo.y, vals[r][2] = lo.z, vals[r][3] = lo.w;
                    vals[r][4] = hi4.x, vals[r][5] = hi4.y, vals[r][6] = hi4.z;
                    acc_s[(r * 2 + 0) * THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
                    acc_s[(r * 2 + 1) * THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        vals[r][q] = acc[r][q];
                        acc[r][q] = 0.f;
                    }
                }
            }
            bool finalize = complete;
            if (!complete) {
                // the plane is shared with neighbouring CTAs: park the partial sums, the last to arrive adds
                // the parts in CTA order.  Slot 0 = the plane this CTA's run starts in, slot 1 = any other.
                float *mine = p.scratch + ((size_t)blockIdx.x * 2 + (vp == vp_lo ? 0 : 1)) * 7 * hm;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int k = slot_gate(r);
#pragma unroll
                    for (int q = 0; q < 7; ++q) mine[q * hm + k] = vals[r][q];
                    WRP_CHECK(k >= 0 && k < hm && blockIdx.x < gridDim.x);
                }
                __threadfence();
                __syncthreads();
                WRP_CHECK(plane >= 0 && plane < p.S * p.C);
                if (tid == 0) s_flag = atomicAdd(p.plane_cnt + plane, 1);
                __syncthreads();
                const int x_first = cta_of(plane_first), x_last = cta_of(plane_first + p.NT - 1);
                if (s_flag == x_last - x_first) {
                    __threadfence();
#pragma unroll
                    for (int r = 0; r < RPT; ++r)
#pragma unroll
                        for (int q = 0; q < 7; ++q) vals[r][q] = 0.f;
                    // every CTA after the first starts its run inside this plane (slot 0); the first one only if its
                    // run begins exactly at the plane's first tile.  One sector alone is cut into NT parts per plane,
                    // so this loop is the tail of the single-sector latency: no division in it, four parts' loads in
                    // flight, the additions still in CTA order.  (Unrolled further, slot by slot with eight parts in
                    // flight, it is no faster and the larger kernel image costs the full-batch rate 2 %.)
                    const int first_slot = (int)(((long long)x_first * total / Gp) / p.NT) == vp ? 0 : 1;
#pragma unroll 4
                    for (int x = x_first; x <= x_last; ++x) {
                        const float *part =
                            p.scratch + ((size_t)(x * CG + grp) * 2 + (x == x_first ? first_slot : 0)) * 7 * hm;
#pragma unroll
                        for (int r = 0; r < RPT; ++r) {
                            const int k = slot_gate(r);
#pragma unroll
                            for (int q = 0; q < 7; ++q) vals[r][q] += __ldcg(part + q * hm + k);
                        }
                    }
                    finalize = true;
                }
            }
            if (finalize) {
                const int sector = plane / p.C, ch = plane - sector * p.C;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int k = slot_gate(r);
                    float removed = vals[r][1] * vals[r][1];
#pragma unroll
                    for (int q = 2; q < 7; ++q) removed = fmaf(vals[r][q], vals[r][q], removed);
                    // stages 05-08: the row sum of the circular convolution is sum(taps) x the row sum
                    const float pw = fmaxf(fmaf(p.n_float, vals[r][0], -removed), 0.f) * p.taps_sum;
                    WRP_CHECK(plane >= 0 && plane < p.S * p.C && k >= 0 && k < hm && sector >= 0 && sector < p.S);
                    p.power[(size_t)plane * hm + k] = pw;
                    if (p.C == 1) { // stage 09 only (rpv2.cu:199-213 with a single channel)
                        const float rg = (float)k * p.range_res;
                        store_product(p, (size_t)sector * hm + k, make_float2(10.f * log10f(rg * rg * p.calib * pw), 0.f));
                    }
                }
                if (p.C >= 2 && ch < 2) {
                    // stages 09/10 need hh and vv of the gate: whoever finishes the second of the two planes
                    __threadfence();
                    __syncthreads();
                    if (tid == 0) s_flag = atomicAdd(p.sector_cnt + sector, 1);
                    __syncthreads();
                    if (s_flag == 1) {
                        __threadfence();
                        const float *ph = p.power + (size_t)sector * p.C * hm, *pv = ph + hm;
                        for (int k = tid; k < hm; k += THREADS) {
                            const float hh = __ldcg(ph + k), vv = __ldcg(pv + k);
                            const float rg = (float)k * p.range_res;
                            store_product(p, (size_t)sector * hm + k,
                                          make_float2(10.f * log10f(rg * rg * p.calib * hh), 10.f * (log10f(hh) - log10f(vv))));
                        }
                    }
                }
            }
        }
        first = false;
        t = nt;
        vp = nvp;
    }
}


// ================================================================================================
// chain_wire3_kernel — wire input (12-byte records hhI hhQ vvI vvQ vhI vhQ, sector.cpp:52-62), M = 1024,
// three channels, ALL of them in one CTA.
//
// A tile = 4 adjacent record columns x 1024 sweeps = 12 FFT columns (4 columns x hh, vv, vh).  Its raw
// rows (48 bytes) arrive by TMA — four boxes {12 x uint32, 256 rows} — into one of two 48 KiB landing
// buffers, separate from the exchange buffer, so the load of tile n+1 is issued at the top of tile n and has
// the whole tile's work to land (chain_stream_kernel<1, true> gathers one channel with 4-byte cp.async, which
// the LSU handles element by element: 0.74 TB/s, tools/micro/tile_load.cu; TMA cannot gather below 16 bytes,
// raw rows come in at 4.3 TB/s).  384 threads, one CTA per SM:
//   pass 1  thread (c2, b), c2 = record column * 3 + channel = position of the (I, Q) pair in the raw row:
//           rows 32 a + b, big-endian decode, window in the first butterfly stage, radix-32, twiddle,
//           exchange buffer [32 ka groups][32 rows][96 B] with 96 B of padding per group (conflict-free)
//   pass 2  warp = (channel, 8 ka), lane = (ka, column): radix-32 over b, outputs k < 512
//   fold    as chain_stream_kernel's M = 4096 form (4-column rows, four rounds, warp-local staging); a thread
//           owns four (gate, channel) rows: 28 accumulators in registers
//   sector end: P[k] per channel -> power[], ZdB/ZDR from hh and vv by the same CTA (rpv2.cu:199-213); a sector
//           cut by the work partition is summed by the last CTA to arrive, parts in CTA order.
// No CTA waits for another.
namespace w3 {
constexpr int THREADS = 384, NWARP = 12, COLS = 12, RC = 4; // record columns per tile
constexpr int RAW_PITCH = 48, RAW_BYTES = 1024 * RAW_PITCH;
constexpr int XROW = 96, XGROUP = 32 * XROW + 96, XBUF = 32 * XGROUP;
constexpr int STAGE_WARP = 32 * 32, STAGE = NWARP * STAGE_WARP;
constexpr int OFF_RAW = 0, OFF_X = 2 * RAW_BYTES, OFF_STAGE = OFF_X + XBUF, OFF_WRC = OFF_STAGE + STAGE;
constexpr int OFF_TWA = OFF_WRC + 32 * WRC_ROW, OFF_BAR = OFF_TWA + 32 * TWA_ROW, SMEM = OFF_BAR + 64;
static_assert(SMEM <= 232448, "shared memory per CTA");
} // namespace w3

__global__ void __launch_bounds__(w3::THREADS, 1)
    chain_wire3_kernel(const StreamParams p, const __grid_constant__ CUtensorMap tmap)
{
    using namespace w3;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *const rbar = reinterpret_cast<uint64_t *>(smem + OFF_BAR);     // [2] raw tile landed
    uint64_t *const ebar = reinterpret_cast<uint64_t *>(smem + OFF_BAR + 16); // every warp has read its pass-2 operands
    int *const s_flag = reinterpret_cast<int *>(smem + OFF_BAR + 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hm = 512;

    // work partition: a contiguous run of tiles of the [sector][tile] space
    const int NT = p.N / RC;
    const long long total = (long long)p.S * NT;
    const int G = gridDim.x;
    const int g_lo = (int)((long long)blockIdx.x * total / G), g_end = (int)(((long long)blockIdx.x + 1) * total / G);
    if (g_lo >= g_end) return;
    auto cta_of = [&](long long g) { return (int)(((g + 1) * G - 1) / total); };

    for (int i = tid; i < 32 * 8; i += THREADS) // 16-byte pieces of the window table
        *reinterpret_cast<float4 *>(smem + OFF_WRC + (i >> 3) * WRC_ROW + (i & 7) * 16) =
            __ldg(reinterpret_cast<const float4 *>(p.wrc_t) + i);
    for (int i = tid; i < 32 * 16; i += THREADS)
        *reinterpret_cast<float4 *>(smem + OFF_TWA + (i >> 4) * TWA_ROW + (i & 15) * 16) =
            __ldg(reinterpret_cast<const float4 *>(p.tw_a) + i);
    if (tid == 0) {
        mbar_init(&rbar[0], 1);
        mbar_init(&rbar[1], 1);
        mbar_init(ebar, NWARP);
    }
    __syncthreads();

    int sector = g_lo / NT, t = g_lo - sector * NT;
    const int sector_lo = sector;
    auto issue_raw = [&](int sec, int tile, int buf) { // thread 0: four boxes of 256 sweeps x 48 bytes
        mbar_expect_tx(&rbar[buf], RAW_BYTES);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            tma_load_2d(smem + OFF_RAW + buf * RAW_BYTES + q * (256 * RAW_PITCH), &tmap, tile * COLS, sec * 1024 + q * 256,
                        &rbar[buf]);
    };
    if (tid == 0) issue_raw(sector, t, 0);

    // pass-1 identity: c2 = position of the thread's (I, Q) pair in the raw row, b = row residue
    const int c2 = tid % COLS, b1 = tid / COLS;
    // pass-2 / fold identity: warp = (channel, group of eight ka), lane = (ka, column)
    const int ch = warp >> 2, kag = warp & 3, ka_l = lane >> 2, colw = lane & 3;
    const int ka = 8 * kag + ka_l;
    uint8_t *const stage = smem + OFF_STAGE + warp * STAGE_WARP;
    // staged row R = kbl * 8 + ka_l (32-byte rows); 16-byte chunk (col >> 1) XOR s(R), s(R) = (R >> 2) & 1
    uint8_t *const wbase = stage + ka_l * 32 + ((((colw >> 1) ^ ((ka_l >> 2) & 1))) << 4) + (colw & 1) * 8;
    const int rs = (lane >> 2) & 1;
    const uint8_t *const rbase = stage + lane * 32;
    auto slot_gate = [&](int r) { return 8 * kag + (lane & 7) + 32 * (4 * r + (lane >> 3)); };
    float acc[4][7];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[r][q] = 0.f;

    uint32_t ephase = 0;
    int n = 0;
    for (int g = g_lo; g < g_end; ++g, ++n) {
        int nt = t + 1, nsector = sector;
        if (nt == NT) nt = 0, ++nsector;
        const bool has_next = g + 1 < g_end;
        const int buf = n & 1;
        // the other landing buffer was last read in pass 1 of the previous tile, and every thread has passed
        // that tile's exchange barrier: request the next tile now, a whole tile ahead
        if (tid == 0 && has_next) issue_raw(nsector, nt, buf ^ 1);
        WRP_CHECK(t >= 0 && t < NT && sector >= 0 && sector < p.S && g < total);
        const float4 ttw = __ldg(p.tile_tw + t);
        const float wdj = __ldg(p.wd + t * RC + c2 / 3);

        mbar_wait(&rbar[buf], (n >> 1) & 1);

        // ================= pass 1 =================
        float2 v[R];
        {
            const uint8_t *src = smem + OFF_RAW + buf * RAW_BYTES + b1 * RAW_PITCH + c2 * 4;
            static_for<R>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                const uint32_t w = *reinterpret_cast<const uint32_t *>(src + a * (R * RAW_PITCH));
                v[brev<R>(a)] = make_float2((float)(int)prmt(w, 0u, 0x8801u), (float)(int)prmt(w, 0u, 0xAA23u));
            });
        }
        {
            const float2 m2 = make_float2(-2.f, -2.f);
            const float4 *w4 = reinterpret_cast<const float4 *>(smem + OFF_WRC + b1 * WRC_ROW);
            static_for<R / 8>([&](auto qi) {
                constexpr int q = decltype(qi)::value;
                const float4 wlo = w4[q], whi = w4[q + R / 8];
                const float lo[4] = {wlo.x, wlo.y, wlo.z, wlo.w}, hi[4] = {whi.x, whi.y, whi.z, whi.w};
                static_for<4>([&](auto ei) {
                    constexpr int e = decltype(ei)::value;
                    constexpr int sa = brev<R>(4 * q + e);
                    const float wl = lo[e] * wdj, wh = hi[e] * wdj;
                    const float2 tt = cmul2(v[sa + 1], make_float2(wh, wh));
                    const float2 s2 = cfma2(v[sa], make_float2(wl, wl), tt);
                    v[sa + 1] = cfma2(tt, m2, s2);
                    v[sa] = s2;
                });
            });
            fft_dit_after_stage1<R, -1>(v);
        }
        if (n > 0) { // every warp has pulled its pass-2 operands of the previous tile out of the exchange buffer
            mbar_wait(ebar, ephase);
            ephase ^= 1;
        }
        {
            const float4 *t4 = reinterpret_cast<const float4 *>(smem + OFF_TWA + b1 * TWA_ROW);
            uint8_t *dst = smem + OFF_X + b1 * XROW + c2 * 8;
            float4 wq[3] = {t4[0], t4[1], t4[2]};
            static_for<R / 2>([&](auto qi) {
                constexpr int q = decltype(qi)::value;
                const float4 w = wq[q % 3];
                if constexpr (q + 3 < R / 2) wq[q % 3] = t4[q + 3];
                const float2 y0 = q == 0 ? v[0] : cmul(v[2 * q], make_float2(w.x, w.y));
                const float2 y1 = cmul(v[2 * q + 1], make_float2(w.z, w.w));
                *reinterpret_cast<float2 *>(dst + (2 * q) * XGROUP) = y0;
                *reinterpret_cast<float2 *>(dst + (2 * q + 1) * XGROUP) = y1;
            });
        }
        __syncthreads(); // the exchange

        // ================= pass 2 =================
        {
            const uint8_t *src = smem + OFF_X + ka * XGROUP + (colw * 3 + ch) * 8;
            static_for<R>([&](auto bi) {
                constexpr int bb = decltype(bi)::value;
                v[brev<R>(bb)] = *reinterpret_cast<const float2 *>(src + bb * XROW);
            });
        }
        __syncwarp();
        dit_stage<R, 1, -1>(v); // consumes every loaded value: the arrive below follows the completion of the reads
        if (lane == 0) mbar_arrive_after(ebar, __float_as_uint(v[R - 1].x) & (uint32_t)p.zero);
        fft_dit_after_stage1<R, -1>(v);
        const int col = t * RC + colw;
        if (p.x2_tap) {
            WRP_CHECK(ka + R * 15 < hm && col >= 0 && col < p.N);
            float2 *o = p.x2_tap + ((size_t)(sector * 3 + ch) * hm + ka) * (size_t)p.N + col;
            static_for<R / 2>([&](auto ki) {
                constexpr int kb = decltype(ki)::value;
                o[(size_t)(R * kb) * p.N] = v[kb];
            });
        }

        // ================= fold =================
        static_for<4>([&](auto hi_) {
            constexpr int h = decltype(hi_)::value;
            static_for<4>([&](auto ki) {
                constexpr int kbl = decltype(ki)::value;
                *reinterpret_cast<float2 *>(wbase + kbl * (8 * 32)) = v[4 * h + kbl];
            });
            __syncwarp();
            {
                float2 x[4];
                static_for<2>([&](auto qi) {
                    constexpr int q = decltype(qi)::value;
                    const float4 w = *reinterpret_cast<const float4 *>(rbase + ((q ^ rs) << 4));
                    x[2 * q] = make_float2(w.x, w.y);
                    x[2 * q + 1] = make_float2(w.z, w.w);
                });
                fold_row<4>(x, p.wcol, ttw, acc[h]);
            }
            __syncwarp();
        });

        // ================= end of the sector (or of this CTA's run) =================
        if (nt == 0 || !has_next) {
            const long long sector_first = (long long)sector * NT;
            const bool complete = g_lo <= sector_first && nt == 0;
            float vals[4][7];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    vals[r][q] = acc[r][q];
                    acc[r][q] = 0.f;
                }
            bool finalize = complete;
            if (!complete) {
                float *mine = p.scratch + (((size_t)blockIdx.x * 2 + (sector == sector_lo ? 0 : 1)) * 3 + ch) * 7 * hm;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int k = slot_gate(r);
#pragma unroll
                    for (int q = 0; q < 7; ++q) mine[q * hm + k] = vals[r][q];
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) *s_flag = atomicAdd(p.plane_cnt + sector, 1);
                __syncthreads();
                const int x_first = cta_of(sector_first), x_last = cta_of(sector_first + NT - 1);
                if (*s_flag == x_last - x_first) {
                    __threadfence();
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int q = 0; q < 7; ++q) vals[r][q] = 0.f;
                    // (slot 0 for every CTA after the first: its run starts inside this sector — see chain_stream_kernel)
                    const int first_slot = (int)(((long long)x_first * total / G) / NT) == sector ? 0 : 1;
#pragma unroll 2
                    for (int x = x_first; x <= x_last; ++x) {
                        const float *part = p.scratch + (((size_t)x * 2 + (x == x_first ? first_slot : 0)) * 3 + ch) * 7 * hm;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int k = slot_gate(r);
#pragma unroll
                            for (int q = 0; q < 7; ++q) vals[r][q] += __ldcg(part + q * hm + k);
                        }
                    }
                    finalize = true;
                }
            }
            if (finalize) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int k = slot_gate(r);
                    float removed = vals[r][1] * vals[r][1];
#pragma unroll
                    for (int q = 2; q < 7; ++q) removed = fmaf(vals[r][q], vals[r][q], removed);
                    WRP_CHECK(k >= 0 && k < hm && ch >= 0 && ch < 3);
                    p.power[((size_t)sector * 3 + ch) * hm + k] =
                        fmaxf(fmaf(p.n_float, vals[r][0], -removed), 0.f) * p.taps_sum;
                }
                __syncthreads(); // hh and vv of a gate were written by different warps of this CTA
                const float *ph = p.power + (size_t)sector * 3 * hm, *pv = ph + hm;
                for (int k = tid; k < hm; k += THREADS) {
                    const float hh = ph[k], vv = pv[k];
                    const float rg = (float)k * p.range_res;
                    store_product(p, (size_t)sector * hm + k,
                                  make_float2(10.f * log10f(rg * rg * p.calib * hh), 10.f * (log10f(hh) - log10f(vv))));
                }
            }
        }
        t = nt;
        sector = nsector;
    }
}

} // namespace stream

// ---- host side ---------------------------------------------------------------------------------
bool stream_supported(int M, int N, int wire)
{
    if (N < 64 || N > 8192 || (N & (N - 1))) return false;
    if (M == 1024) return true;
    return M == 4096 && !wire;
}

const char *stream_kernel_name() { return "chain_stream_kernel"; }

template <int Q, bool WIRE> static cudaError_t setup_one(int sm_count, int *max_grid)
{
    using K = stream::Cfg<Q, WIRE>;
    cudaError_t e = cudaFuncSetAttribute(stream::chain_stream_kernel<Q, WIRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream::chain_stream_kernel<Q, WIRE>, K::THREADS, K::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *max_grid = per_sm * sm_count;
    return cudaSuccess;
}

cudaError_t wire3_setup(int sm_count, int *max_grid)
{
    cudaError_t e = cudaFuncSetAttribute(stream::chain_wire3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stream::w3::SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stream::chain_wire3_kernel, stream::w3::THREADS, stream::w3::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *max_grid = per_sm * sm_count;
    return cudaSuccess;
}

size_t wire3_scratch_floats(int max_grid) { return (size_t)max_grid * 2 * 3 * 7 * 512; }

bool wire3_encode_tensor_map(void *encode_fn, CUtensorMap *out, const void *base, int N, long long sectors)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    if (!encode_fn) return false;
    // the batch as a [sectors * 1024 sweeps][3 N] matrix of 32-bit (I, Q) pairs; box = 12 pairs (4 records) x 256 sweeps
    const cuuint64_t dims[2] = {(cuuint64_t)N * 3, (cuuint64_t)sectors * 1024};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 12};
    const cuuint32_t box[2] = {12, 256};
    const cuuint32_t estr[2] = {1, 1};
    return ((EncodeFn)encode_fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Grid of one launch.  Large batches take every CTA slot.  Small batches are latency problems: a CTA's tiles run back
// to back (about 4 us each) and a plane cut into k parts costs its last arriver k partial-sum reads, so fewer CTAs with
// more tiles each finish earlier — measured on B200 (profiles/r02_small_batch_grid.md): at least min_tiles tiles per CTA
// (4 for the plane kernels, 8 for chain_wire3_kernel), and one CTA per SM as long as the full grid would hold fewer than
// 6 tiles each.
static long long pick_grid(long long total, long long max_grid, int sm_count, int min_tiles = 4)
{
    const long long g = max_grid < total ? max_grid : total;
    const long long g4 = total / min_tiles > 0 ? total / min_tiles : 1;
    if (sm_count < 1) sm_count = 1;
    if (g4 <= sm_count) return g4 < g ? g4 : g;
    if (total < 6 * max_grid && sm_count < g) return sm_count;
    return g;
}

cudaError_t launch_wire3(StreamParams p, int max_grid, int sm_count, const CUtensorMap &tmap, cudaStream_t st)
{
    if (p.S <= 0) return cudaSuccess;
    p.NT = p.N / 4;
    p.half_m = 512;
    p.n_float = (float)p.N;
    p.chan_groups = 1;
    const long long total = (long long)p.S * p.NT;
    const int grid = (int)pick_grid(total, max_grid, sm_count, 8); // 12-column tiles, whole sectors combined by one CTA: 8 per CTA
    cudaError_t e = cudaMemsetAsync(p.plane_cnt, 0, sizeof(int) * (size_t)p.S, st); // parts of a cut sector that have arrived
    if (e != cudaSuccess) return e;
    stream::chain_wire3_kernel<<<grid, stream::w3::THREADS, stream::w3::SMEM, st>>>(p, tmap);
    return cudaGetLastError();
}

cudaError_t stream_setup(int M, int wire, int sm_count, int *max_grid)
{
    if (M == 4096) return setup_one<4, false>(sm_count, max_grid);
    return wire ? setup_one<1, true>(sm_count, max_grid) : setup_one<1, false>(sm_count, max_grid);
}

size_t stream_scratch_floats(int M, int max_grid) { return (size_t)max_grid * 2 * 7 * (M / 2); }

bool stream_encode_tensor_map(void *encode_fn, CUtensorMap *out, const void *base, int M, int N, long long planes,
                              int l2_promotion_bytes)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    if (!encode_fn) return false;
    const int T = M == 4096 ? 4 : 8;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)planes * M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 8};
    const cuuint32_t box[2] = {(cuuint32_t)T, (cuuint32_t)(8192 / (8 * T))};
    const cuuint32_t estr[2] = {1, 1};
    // complex floats travel as opaque 8-byte elements
    return ((EncodeFn)encode_fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 l2_promotion_bytes >= 256   ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                 : l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                 : l2_promotion_bytes >= 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                             : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch_stream(StreamParams p, int M, int wire, int max_grid, int sm_count, int channel_groups,
                          const CUtensorMap &tmap, cudaStream_t st)
{
    if (p.S <= 0) return cudaSuccess;
    const int T = M == 4096 ? 4 : 8;
    p.NT = p.N / T;
    p.half_m = M / 2;
    p.n_float = (float)p.N;
    p.chan_groups = (wire || channel_groups) ? p.C : 1;
    const long long total = (long long)(p.chan_groups == 1 ? p.S * p.C : p.S) * p.NT;
    long long grid = pick_grid(total, max_grid / p.chan_groups, sm_count / p.chan_groups);
    grid *= p.chan_groups;
    cudaError_t e = cudaMemsetAsync(p.plane_cnt, 0, sizeof(int) * (size_t)p.S * (p.C + 1), st); // plane_cnt + sector_cnt
    if (e != cudaSuccess) return e;
    if (M == 4096)
        stream::chain_stream_kernel<4, false><<<(int)grid, stream::Cfg<4, false>::THREADS, stream::Cfg<4, false>::SMEM, st>>>(p, tmap);
    else if (wire)
        stream::chain_stream_kernel<1, true><<<(int)grid, stream::Cfg<1, true>::THREADS, stream::Cfg<1, true>::SMEM, st>>>(p, tmap);
    else
        stream::chain_stream_kernel<1, false><<<(int)grid, stream::Cfg<1, false>::THREADS, stream::Cfg<1, false>::SMEM, st>>>(p, tmap);
    return cudaGetLastError();
}

} // namespace wrp
