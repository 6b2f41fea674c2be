// ptx.cuh — the few sm_100a PTX primitives the scan kernels use directly: mbarrier, the 1-D bulk
// asynchronous copy (cp.async.bulk, SASS UBLKCP — the TMA engine without a tensor map) and the
// proxy fence that orders generic-proxy reads of a shared-memory slot before the next bulk write.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ugx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// arrive (1 of the barrier's count) and announce `bytes` of pending bulk-copy traffic
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity))
  {
  }
}

// global -> shared bulk copy; dst, src and bytes are multiples of 16; completion is signalled on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// read-only table lookups by shared-space address (LDS with a 32-bit address; a generic pointer that may be shared or
// global compiles to LD.E with 64-bit address arithmetic).  Not volatile: the tables are constant once staged.
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
  uint32_t v;
  asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr)
{
  uint32_t v;
  asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

__device__ __forceinline__ uint4 lds128(const void* p)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}

__device__ __forceinline__ uint2 lds64(const void* p)
{
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)));
  return v;
}

// Stage a pattern's tables in shared memory with the bulk-copy engine: one thread issues up to four 1-D copies
// against one mbarrier, every thread of the CTA waits on it.  All sizes are multiples of 16 bytes, all pointers
// 16-byte aligned.  Call from every thread of the CTA, once per kernel (phase 0 of `bar`).
__device__ __forceinline__ void stage_tables_bulk(uint64_t* bar, void* s_cls, const void* g_cls, void* s_pred, const void* g_pred,
                                                  void* s_tap, const void* g_tap, void* s_next, const void* g_next,
                                                  uint32_t next_bytes)
{
  if (threadIdx.x == 0)
  {
    mbar_init(bar, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    mbar_arrive_expect_tx(bar, 256 + 4096 + 2048 + next_bytes);
    bulk_copy_g2s(s_cls, g_cls, 256, bar);
    bulk_copy_g2s(s_pred, g_pred, 4096, bar);
    bulk_copy_g2s(s_tap, g_tap, 2048, bar);
    if (next_bytes != 0)
      bulk_copy_g2s(s_next, g_next, next_bytes, bar);
  }
  mbar_wait(bar, 0);
  __syncthreads();
}

} // namespace ugx
