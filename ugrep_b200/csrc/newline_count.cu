// newline_count.cu — reflex::nlcount (lib/simd.cpp:62-166, the AVX-512BW / AVX2 / SSE2 / NEON newline counters behind
// AbstractMatcher::lineno(), absmatcher.h:695-766) as a streaming kernel: 16 bytes per lane per load, an exact SWAR
// "byte == '\n'" test per 32-bit word, per-byte-lane counters folded once per thread.  HBM-bound.
#include "scan_kernels.hpp"
#include "stream_common.cuh"

namespace ugx {

__global__ void __launch_bounds__(256)
count_newlines_kernel(const uint8_t* __restrict__ buf, uint64_t n, unsigned long long* __restrict__ total)
{
  const uint64_t nvec = n / 16;
  const uint4* __restrict__ p = reinterpret_cast<const uint4*>(buf);
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  unsigned long long count = 0;
  uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  while (i < nvec)
  {
    // up to 31 vectors (124 words) per round keep the four byte-lane counters below 256
    uint32_t acc = 0;
    for (int k = 0; k < 31 && i < nvec; ++k, i += stride)
    {
      const uint4 v = __ldg(p + i);
      acc += zero_bytes(v.x ^ 0x0a0a0a0au) >> 7;
      acc += zero_bytes(v.y ^ 0x0a0a0a0au) >> 7;
      acc += zero_bytes(v.z ^ 0x0a0a0a0au) >> 7;
      acc += zero_bytes(v.w ^ 0x0a0a0a0au) >> 7;
    }
    const uint32_t pair = (acc & 0x00ff00ffu) + ((acc >> 8) & 0x00ff00ffu);
    count += (pair & 0xffffu) + (pair >> 16);
  }
  // the last n % 16 bytes
  if (blockIdx.x == 0 && threadIdx.x < (n & 15))
    count += __ldg(buf + nvec * 16 + threadIdx.x) == '\n';
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    count += __shfl_down_sync(0xffffffffu, count, d);
  __shared__ unsigned long long s_sum[8];
  if ((threadIdx.x & 31) == 0)
    s_sum[threadIdx.x >> 5] = count;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned long long x = 0;
    for (int w = 0; w < 8; ++w)
      x += s_sum[w];
    atomicAdd(total, x);
  }
}

cudaError_t launch_count_newlines(const uint8_t* buf, uint64_t n, unsigned long long* total, int sm_count, cudaStream_t st)
{
  cudaError_t e = cudaMemsetAsync(total, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess)
    return e;
  uint64_t g = (n / 16 + 255) / 256;
  if (g > static_cast<uint64_t>(sm_count) * 8)
    g = static_cast<uint64_t>(sm_count) * 8;
  if (g == 0)
    g = 1;
  count_newlines_kernel<<<static_cast<int>(g), 256, 0, st>>>(buf, n, total);
  return cudaGetLastError();
}

} // namespace ugx
