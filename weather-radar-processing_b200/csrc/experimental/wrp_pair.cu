// wrp_pair.cu — EXPERIMENTAL, NOT PART OF libwrp.so AND NOT YET VALIDATED ON A GPU.
//
// Round-2 candidate for the default sector shape (M = 1024, N = 512), written at the end of round 1
// when the GPU budget was spent: it compiles for sm_100a and passes tools/check_sass.py, nothing
// more.  `make pair` links it INSTEAD of wrp_unified.o into tools/libwrp_pair.so (same three entry
// points: unified_supported / unified_setup / launch_unified); run the parity tests and
// tools/ab.py with WRP_LIB=tools/libwrp_pair.so before believing anything about it.
//
// Why: ncu on chain_unified_kernel (profiles/r01_chain_hot_sass.txt) shows the LSU data pipe as the
// busiest unit (66 %), and the tile loads paying 8 shared-memory wavefronts per cp.async
// instruction instead of 4 — an 8-column tile row is 64 B, half an L2 line, and the shared-memory
// side of cp.async spends one wavefront per returned line.  Here one 16-warp CTA per SM runs two
// 8-warp groups; a work item is TWO adjacent column tiles (16 columns = 128-byte row segments, whole
// lines) plus sixteen Doppler rows.  Every cp.async instruction fetches 4 rows x 128 B and splits
// each line between the two groups' tile buffers (the second buffer is shifted by 64 B so that the
// halves of a line fall into complementary banks).  Each group runs the unified kernel's item body
// on its own 8 columns with its own named barrier; the groups meet only
//   * warp-pairwise (bar.sync 3 + w, 64 threads) before the next tile is requested — warp (g, w)
//     fetches rows [128 w + 64 g, +64) of BOTH buffers, so both warps must be done reading them;
//   * on the wait path of an unmet dependency (__syncthreads).
// Thread 0 decides for the whole CTA (items, dependency bits); group 1 reads the decision after the
// pairwise barrier, which orders it behind thread 0's publication.
//
// Shared memory: 2 x 64 KiB tiles + 64 B shift + 64 KiB rows + tables (window table duplicated
// again for FMUL2) = 211.6 KiB, one CTA per SM.
#include <cstdlib>
#include <cstring>

#include "../wrp_chain_params.h"
#include "../wrp_fft.cuh"
#include "../wrp_internal.h"
#include "../wrp_ptx.cuh"

namespace wrp {

namespace pr {
constexpr int N = 512, R1B = 16, T = 8, NW = 16, GT = 256, THREADS = 512, R = 32;
constexpr int PITCH = T * 8;              // bytes per row of one group's tile
constexpr int OFF_TILE1 = 65536 + 64;     // group 1's tile: shifted by half a bank line
constexpr int OFF_ROWS = 131072 + 128;    // Doppler rows, 4 KiB per warp
constexpr int WRC_ROW = 32 * 8 + 16;      // wr(i)*c transposed [32 b][32 a], each value twice (w, w)
constexpr int TWA_ROW = 32 * 8 + 16;      // range inter-pass twiddles [32 b][32 ka] float2
constexpr int OFF_WRC = OFF_ROWS + NW * 4096;
constexpr int OFF_TWA = OFF_WRC + 32 * WRC_ROW;
constexpr int OFF_WD = OFF_TWA + 32 * TWA_ROW;
constexpr int OFF_TL = OFF_WD + N * 4;
constexpr int SMEM = OFF_TL + 32 * 16;
static_assert(SMEM + 2048 <= 232448, "one CTA per SM, 227 KiB");

struct Item {
    int sa;   // sector of the tile pair (the Doppler rows belong to sector sa - lag); < 0: queue empty
    int sub2; // tile-pair / row-group-pair index inside the sector
    int slot_a, slot_b;
};
} // namespace pr

using pr::Item;

__device__ __forceinline__ Item pair_decode(int idx, const PersistParams &p, int ta2)
{
    Item it;
    it.sa = idx / ta2;
    it.sub2 = idx - it.sa * ta2;
    it.slot_a = it.sa % p.ring;
    it.slot_b = it.sa >= p.lag ? (it.sa - p.lag) % p.ring : 0;
    return it;
}

// x2-ring row of warp gw in row group `sub` (as in wrp_unified.cu)
__device__ __forceinline__ const uint8_t *pair_row(const PersistParams &p, int slot, int sub, int pair_groups, int gw,
                                                   int &chn, int &gate)
{
    if (sub < pair_groups) {
        chn = gw & 1;
        gate = sub * 4 + (gw >> 1);
    } else {
        chn = p.C == 1 ? 0 : 2;
        gate = (sub - pair_groups) * 8 + gw;
    }
    return (const uint8_t *)p.x2 + (((size_t)slot * p.C + chn) * p.half_m + gate) * (size_t)(pr::N * 8);
}

__global__ void __launch_bounds__(pr::THREADS, 1) chain_pair_kernel(const PersistParams p)
{
    using namespace pr;
    constexpr int SW = 128 / PITCH - 1; // row-swizzle mask of the in-place exchange (= 1)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar; // both tiles of the item have landed (one arrival per thread)
    __shared__ int4 s_item[2];
    __shared__ int s_go[2];
    __shared__ float p_row[2][NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp >> 3, gw = warp & 7, gtid = tid & (GT - 1);
    uint8_t *const tile = smem + (grp ? OFF_TILE1 : 0); // this group's 8 columns
    const int TA2 = 32 * p.C;                    // tile pairs per sector
    const int TGT = 64 * p.C;                    // a_done / b_done: one count per tile / row group
    const int pair_groups = p.C >= 2 ? 128 : 0;
    const int total = (p.S + p.lag) * TA2;
    int *const a_done = p.ctrl + CTRL_A, *const b_done = p.ctrl + CTRL_A + p.smax;

    for (int i = tid; i < 32 * 32; i += THREADS) {
        const float w = __ldg(p.wrc_t + i);
        *reinterpret_cast<float2 *>(smem + OFF_WRC + (i >> 5) * WRC_ROW + (i & 31) * 8) = make_float2(w, w);
        *reinterpret_cast<float2 *>(smem + OFF_TWA + (i >> 5) * TWA_ROW + (i & 31) * 8) = __ldg(p.tw_a + i);
    }
    for (int i = tid; i < N; i += THREADS) reinterpret_cast<float *>(smem + OFF_WD)[i] = __ldg(p.wd + i);
    if (tid < 32) {
        float s1, c1, s2, c2;
        sincospif(-2.f * (float)tid / (float)N, &s1, &c1);
        sincospif(-4.f * (float)tid / (float)N, &s2, &c2);
        const float sg = (tid & 1) ? -1.f : 1.f;
        *reinterpret_cast<float4 *>(smem + OFF_TL + tid * 16) = make_float4(sg * c1, sg * s1, sg * c2, sg * s2);
    }

    auto dep_a = [&](const Item &x) -> const int * {
        return (x.sa < p.S && x.sa >= p.ring) ? b_done + (x.sa - p.ring) : nullptr;
    };
    auto dep_b = [&](const Item &x) -> const int * { return (x.sa >= p.lag) ? a_done + (x.sa - p.lag) : nullptr; };

    int next_idx = 0; // thread 0: queue index of the item after the current one (round-robin dealing)
    if (tid == 0) {
        mbar_init(&mbar, THREADS);
        const int first = blockIdx.x;
        next_idx = first + gridDim.x;
        Item f{-1, 0, 0, 0};
        if (first < total) {
            f = pair_decode(first, p, TA2);
            if (const int *d = dep_a(f)) spin_until(d, TGT);
            if (const int *d = dep_b(f)) spin_until(d, TGT);
        }
        s_item[0] = make_int4(f.sa, f.sub2, f.slot_a, f.slot_b);
        s_go[0] = 3;
    }
    __syncthreads();
    Item it{s_item[0].x, s_item[0].y, s_item[0].z, s_item[0].w};

    // two cp.async groups per thread and item, committed in this order: row, tile part
    auto issue_loads_row = [&](const Item &x) {
        if (x.sa >= p.lag) {
            int chn, gate;
            const uint8_t *src = pair_row(p, x.slot_b, 2 * x.sub2 + grp, pair_groups, gw, chn, gate) + lane * 16;
            uint8_t *dst = smem + OFF_ROWS + warp * 4096 + lane * 16;
#pragma unroll
            for (int k = 0; k < 8; ++k) cp_async16(dst + k * 512, src + k * 512);
        }
        cp_async_commit();
    };
    // warp (g, w): rows [128 w + 64 g, +64) x 128 B of the tile pair; chunk q of a row goes to group
    // q / 4's buffer.  One instruction = 4 rows x one whole L2 line each.
    auto issue_loads_tile = [&](const Item &x) {
        if (x.sa < p.S) {
            const int ch = x.sub2 >> 5, col_pair = x.sub2 & 31, q = lane & 7;
            const int row0 = gw * 128 + grp * 64 + (lane >> 3);
            const uint8_t *src = (const uint8_t *)p.iq + ((size_t)(x.sa * p.C + ch) * 1024 + row0) * (N * 8) + col_pair * 128 +
                                 q * 16;
            uint8_t *dst = smem + ((q >> 2) ? OFF_TILE1 : 0) + row0 * PITCH + (q & 3) * 16;
            if (p.evict_first) {
                const uint64_t pol = policy_evict_first();
                const uint32_t d = opaque_smem_addr(dst, p.zero);
#pragma unroll
                for (int k = 0; k < 16; ++k) cp_async16_evict_first(d + k * 4 * PITCH, src + (size_t)k * 4 * (N * 8), pol);
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) cp_async16(dst + k * 4 * PITCH, src + (size_t)k * 4 * (N * 8));
            }
            cp_async_arrive(&mbar);
        }
        cp_async_commit();
    };
    if (it.sa >= 0) {
        issue_loads_row(it);
        issue_loads_tile(it);
    }

    uint32_t phase = 0;
    int n = 0;
    int pending = -1;       // sector of this group's finished tile whose completion is not published yet
    bool rows_late = false; // the current item's rows were requested after its tile (dependency wait)

    while (it.sa >= 0) {
        const bool has_a = it.sa < p.S, has_b = it.sa >= p.lag;
        const int nslot = (n + 1) & 1;
        Item cand{-1, 0, 0, 0};
        const int *pa = nullptr, *pb = nullptr;
        int va = 0, vb = 0, next_idx2 = 0;
        if (tid == 0) {
            if (next_idx < total) {
                cand = pair_decode(next_idx, p, TA2);
                pa = dep_a(cand);
                pb = dep_b(cand);
                if (pa) va = ld_relaxed(pa);
                if (pb) vb = ld_relaxed(pb);
            }
            next_idx2 = next_idx + gridDim.x;
        }
        auto publish_next = [&]() {
            if (tid == 0) {
                bool ok_a = !pa || va >= TGT, ok_b = !pb || vb >= TGT;
                if (!ok_a) ok_a = ld_relaxed(pa) >= TGT;
                if (!ok_b) ok_b = ld_relaxed(pb) >= TGT;
                if ((p.debug & 16) && cand.sa >= 0) {
                    if (!ok_a) atomicAdd(p.ctrl + 1, 1);
                    if (!ok_b) atomicAdd(p.ctrl + 2, 1);
                }
                s_item[nslot] = make_int4(cand.sa, cand.sub2, cand.slot_a, cand.slot_b);
                s_go[nslot] = (ok_a ? 1 : 0) | (ok_b ? 2 : 0);
            }
        };

        // ---- Doppler row of this warp: stages 03-08 in energy form (wrp_persistent.cu, DOP == 1) ----
        float pw = 0.f;
        if (has_b) {
            if (rows_late) cp_async_wait_group<0>();
            else cp_async_wait_group<1>();
            __syncwarp();
            const uint8_t *row = smem + OFF_ROWS + warp * 4096;
            float2 v[R1B];
            static_for<R1B>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                v[a] = *reinterpret_cast<const float2 *>(row + (32 * a + lane) * 8);
            });
            const float4 tl = *reinterpret_cast<const float4 *>(smem + OFF_TL + lane * 16);
            float2 e2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            static_for<R1B>([&](auto ai) {
                constexpr int a = decltype(ai)::value;
                e2[a & 1] = cfma2(v[a], v[a], e2[a & 1]);
            });
            const float2 es = cadd(e2[0], e2[1]);
            float2 b0, b1, b2;
            dft_bins012<R1B>(v, b0, b1, b2);
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
            pw = fmaxf(fmaf((float)N, r[0], -removed), 0.f) * p.taps_sum;
            if (lane == 0) p_row[n & 1][warp] = pw;
        }

        // ---- range tile of this group, first pass ----
        const int c = gtid % T, b = gtid / T;
        const int ch = it.sub2 >> 5, col = (it.sub2 & 31) * 16 + grp * T + c;
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
                const float wdj = reinterpret_cast<const float *>(smem + OFF_WD)[col];
                const float2 wd2 = make_float2(wdj, wdj), m2 = make_float2(-2.f, -2.f);
                const float4 *w4 = reinterpret_cast<const float4 *>(smem + OFF_WRC + b * WRC_ROW);
                static_for<R / 4>([&](auto qi) { // two rows a, a+1 per 128-bit table read
                    constexpr int q = decltype(qi)::value;
                    const float4 wlo = w4[q], whi = w4[q + R / 4];
                    static_for<2>([&](auto ei) {
                        constexpr int e = decltype(ei)::value;
                        constexpr int sa = brev<R>(2 * q + e);
                        static_assert(brev<R>(2 * q + e + R / 2) == sa + 1, "span-1 partner");
                        const float2 wl = cmul2(e ? make_float2(wlo.z, wlo.w) : make_float2(wlo.x, wlo.y), wd2);
                        const float2 wh = cmul2(e ? make_float2(whi.z, whi.w) : make_float2(whi.x, whi.y), wd2);
                        const float2 t = cmul2(v[sa + 1], wh);
                        const float2 s2 = cfma2(v[sa], wl, t);
                        v[sa + 1] = cfma2(t, m2, s2);
                        v[sa] = s2;
                    });
                });
                fft_dit_after_stage1<R, -1>(v);
            }
            publish_next();
            __syncwarp();
            {
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
        bar_sync(1 + grp, GT); // the exchange of this group's eight columns

        // ---- right after the group barrier: this group's publications and products ----
        if (pending >= 0 && gtid == GT - 32) red_release_add(a_done + pending);
        pending = -1;
        if (has_b) {
            const int sb = it.sa - p.lag, sub = 2 * it.sub2 + grp;
            if (gtid == GT - 64) atomicAdd(b_done + sb, 1);
            if (lane == 0) {
                int chn, gate;
                (void)pair_row(p, it.slot_b, sub, pair_groups, gw, chn, gate);
                if (p.power) p.power[((size_t)sb * p.C + chn) * p.half_m + gate] = pw;
                const bool pair = sub < pair_groups;
                if ((pair && !(gw & 1)) || (!pair && p.C == 1)) {
                    const float rg = (float)gate * p.range_res;
                    const float z = rg * rg * p.calib * pw;
                    reinterpret_cast<float2 *>(p.out)[(size_t)sb * p.half_m + gate] =
                        make_float2(10.f * log10f(z), pair ? 10.f * (log10f(pw) - log10f(p_row[n & 1][warp + 1])) : 0.f);
                }
            }
        }

        // ---- second pass reads, then the rendezvous of the two warps that share tile rows ----
        const int ka = b;
        if (has_a) {
            const uint8_t *s_sw[SW + 1];
#pragma unroll
            for (int sx = 0; sx <= SW; ++sx) s_sw[sx] = tile + ka * (R * PITCH) + c * 8 + ((ka ^ sx) & SW) * PITCH;
            static_for<R>([&](auto bi) {
                constexpr int bb = decltype(bi)::value;
                v[brev<R>(bb)] = *reinterpret_cast<const float2 *>(s_sw[bb & SW] + (bb & ~SW) * PITCH);
            });
        }
        bar_sync(3 + gw, 64); // warps (0, gw) and (1, gw): rows [128 gw, +128) of both tiles are in registers
        const Item nit{s_item[nslot].x, s_item[nslot].y, s_item[nslot].z, s_item[nslot].w};
        const int go_bits = nit.sa >= 0 ? s_go[nslot] : 0;
        const bool go_a = go_bits & 1, go_b = go_bits & 2;
        if (go_b) issue_loads_row(nit);
        if (go_a) issue_loads_tile(nit);
        if (has_a) {
            fft_dit<R, -1>(v);
            float2 *out = p.x2 + (((size_t)it.slot_a * p.C + ch) * p.half_m + ka) * (size_t)N + col;
            static_for<R / 2>([&](auto ki) {
                constexpr int kb = decltype(ki)::value;
                out[(size_t)(R * kb) * N] = v[kb];
            });
            pending = it.sa;
        }

        rows_late = false;
        if (nit.sa < 0 || !go_a || !go_b) {
            __syncthreads();
            if (pending >= 0 && gtid == GT - 32) red_release_add(a_done + pending);
            pending = -1;
            if (nit.sa >= 0) {
                if (tid == 0) {
                    if (!go_a)
                        if (const int *d = dep_a(nit)) spin_until(d, TGT);
                    if (!go_b)
                        if (const int *d = dep_b(nit)) spin_until(d, TGT);
                }
                __syncthreads();
                if (!go_b) issue_loads_row(nit);
                if (!go_a) issue_loads_tile(nit);
                rows_late = go_a && !go_b;
            }
        }
        it = nit;
        ++n;
        next_idx = next_idx2;
    }
}

// ---- host side: the three entry points of wrp_unified.cu ---------------------------------------
bool unified_supported(int M, int N) { return M == 1024 && N == 512; }

cudaError_t unified_setup()
{
    return cudaFuncSetAttribute(chain_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pr::SMEM);
}

cudaError_t launch_unified(PersistParams p, int sm_count, cudaStream_t st)
{
    if (p.lag > p.S) p.lag = p.S;
    const int total = (p.S + p.lag) * 32 * p.C;
    int grid = sm_count;
    if (grid > total) grid = total;
    chain_pair_kernel<<<grid, pr::THREADS, pr::SMEM, st>>>(p);
    return cudaGetLastError();
}

} // namespace wrp
