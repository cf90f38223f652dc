// copy_async.cuh -- asynchronous-copy, mbarrier and named-barrier primitives shared by the fused KCF kernels
// (kcf_fused.cuh: fixed power-of-two cell grids; kcf_any.cu: any window size).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mot {

// 8-byte asynchronous copy global -> shared (SASS: LDGSTS), completion tracked per thread with commit / wait groups
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
// barrier among a subset of the CTA's warps (id 1..15; __syncthreads is id 0)
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// the non-blocking half: signal arrival at named barrier `id` (nthreads = arrivers + waiters) and carry on
__device__ __forceinline__ void bar_arrive_named(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// one instruction pulls a whole (16-byte aligned, 16-byte multiple) range into L2
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }

// ---- bulk asynchronous copies global -> shared on an mbarrier (the copy engine that TMA uses; SASS: UBLKCP) --------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 2-D tiled copy through a tensor map (TMA proper; SASS: UTMALDG): the box the map was encoded with, starting at element (c0, c1) of
// the tensor, lands densely in shared memory (128-byte aligned) and completes on the mbarrier with the full box byte count
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tMBW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra MBW_DONE;\n\tbra MBW_LOOP;\n\tMBW_DONE:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

}  // namespace mot
