// wrp_ptx.cuh — PTX wrappers shared by the persistent chain kernels (sm_100a): mbarrier, cp.async
// (plain and with an L2 evict-first hint), named barriers, release/relaxed counters.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace wrp {

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
// 16-byte async copy global -> shared (L1 bypass).  An L2 evict-first cache hint was tried here
// (cp.async ... .L2::cache_hint): ptxas 12.9 allocated an odd uniform register for the LDGSTS
// descriptor at one call site and the warp trapped with "illegal instruction"; st/ld
// eviction-priority qualifiers need 256-bit vectors on sm_100.  WRP_L2_PERSIST pins the ring instead.
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// Same copy with an L2 evict-first policy: the input is streamed once and should not push the x2
// ring out of L2.  (With an earlier code shape ptxas 12.9 gave one call site an ODD uniform
// descriptor register for this instruction form and the warp trapped with "illegal instruction";
// tools/check_sass.py fails the build if that ever comes back.)
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16_evict_first(uint32_t dst_smem, const void *src, uint64_t pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "l"(pol)
                 : "memory");
}
// A shared-memory address ptxas cannot split into register + uniform base (XOR with a kernel
// parameter that is always 0): the hinted LDGSTS has no room for a uniform address offset next to
// its uniform policy operand, and ptxas 12.9 overwrites the low policy word with the offset and
// emits a "[R+UR], desc[URodd]" encoding that traps (tools/check_sass.py guards the build).
__device__ __forceinline__ uint32_t opaque_smem_addr(const void *p, int zero)
{
    return smem_u32(p) ^ (uint32_t)zero;
}
// arrive on `bar` once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// producer/consumer named barrier: warps that only announce do not wait
__device__ __forceinline__ void bar_arrive(int id, int threads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
template <int ID> __device__ __forceinline__ void bar_sync_id(int threads)
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(threads) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// release-add: orders the executing thread's prior writes and, through the preceding __syncwarp,
// its warp's (cumulativity); no L1 invalidation, unlike __threadfence()
__device__ __forceinline__ void red_release_add(int *p)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}
// probe without the L1 invalidate an acquire load carries: the counter is bumped by a release
// (data is in L2 before the count moves) and everything read under it is fetched by cp.async.cg
// straight from L2, issued after the value has been seen
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Bounded: a dependency that never arrives (a kernel bug, a corrupted counter) must end as a device error —
// cudaErrorLaunchFailure -> WRP_ERR_CUDA at the next API call — not as a hung GPU.  2^26 probes x >= 100 ns
// is several seconds, three orders of magnitude beyond the longest legitimate wait (one work item).
__device__ __forceinline__ void spin_until(const int *p, int target)
{
    for (unsigned spins = 0; ld_acquire(p) < target; ++spins) {
        if (spins >> 26) __trap();
        __nanosleep(100);
    }
}
// pairs a successful relaxed probe with the producer's red.release: data read after it (cp.async.cg from L2)
// is ordered behind the counter value in the PTX memory model, not only in practice
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// L2 evict-last policy and an 8-byte store carrying it (experiments on keeping the x2 ring resident)
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_global_hint(float2 *ptr, float2 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(ptr), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

// split-phase CTA rendezvous: arrive now (release), wait later (acquire) — work in between overlaps
// the time the slower warps still need
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte permute with PTX semantics: selector nibble n = byte index 0-7 of {b, a}; bit 3 of a nibble
// replicates that byte's sign bit instead (the __byte_perm intrinsic masks bit 3 away)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// mbarrier arrive that releases a shared-memory region the warp has been READING: `dep_and_zero` = (a value derived
// from the last of those loads) AND (a kernel parameter that is always 0).  Added to the barrier address it changes
// nothing but keeps the arrive behind the COMPLETION of the loads in the SASS, not merely behind their issue.
__device__ __forceinline__ void mbar_arrive_after(uint64_t *bar, uint32_t dep_and_zero)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + dep_and_zero) : "memory");
}

// cp.async groups: the calling thread's copies since its previous commit form one group
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> __device__ __forceinline__ void cp_async_wait_group()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

} // namespace wrp
