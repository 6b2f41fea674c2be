// block_scan.cuh — warp / block exclusive prefix sums used by the line-scan kernels.
#pragma once

#include <cstdint>

namespace ugx {

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane)
{
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= static_cast<uint32_t>(d))
      x += y;
  }
  return x - v;
}

// exclusive scan over the block; total returned through *total (all threads)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total)
{
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t ex = warp_excl_scan(v, lane);
  if (lane == 31)
    warp_sums[wid] = ex + v;
  __syncthreads();
  if (wid == 0)
  {
    uint32_t s = lane < nw ? warp_sums[lane] : 0;
    uint32_t e = warp_excl_scan(s, lane);
    if (lane < nw)
      warp_sums[lane] = e;
    if (lane == 31)
      warp_sums[32] = e + s;
  }
  __syncthreads();
  uint32_t r = ex + warp_sums[wid];
  *total = warp_sums[32];
  __syncthreads();
  return r;
}


} // namespace ugx
