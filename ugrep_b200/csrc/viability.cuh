// viability.cuh — the k-gram viability test of an anchored attempt (pattern_host.hpp Viability): a position whose
// next k <= 4 bytes cannot keep the DFA alive (nor make it accept) starts no match, so the scan kernels spare its
// attempt.  Two shared-memory lookups per position; a property of the DFA alone, it never changes a result.
#pragma once

#include "device_pattern.cuh"
#include "tile_phase_a.cuh"

namespace ugx {

struct ViaTables {
  const uint2* t;       // [256] shared: (t01, t23) of a byte
  const uint8_t* pair;  // shared
  const uint32_t* bits; // shared
  uint32_t stride;
  bool on;
  bool bytes;           // `bits` holds one byte per entry
};

__host__ __device__ inline uint32_t via_smem_bytes(const DevPattern& P) { return P.via_k ? 2048 + P.via_pair_bytes + P.via_words * 4 : 0; }

// stage the tables behind `base` (16-byte aligned); every thread of the CTA calls it; the caller synchronises
__device__ __forceinline__ ViaTables via_stage(const DevPattern& P, uint8_t* base, bool on)
{
  ViaTables v;
  v.on = on && P.via_k != 0;
  v.stride = P.via_stride;
  v.bytes = P.via_bytes != 0;
  uint32_t* t = reinterpret_cast<uint32_t*>(base);
  uint32_t* bits = t + 512;
  uint8_t* pair = reinterpret_cast<uint8_t*>(bits + P.via_words);
  v.t = reinterpret_cast<const uint2*>(t);
  v.bits = bits;
  v.pair = pair;
  if (v.on)
  {
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x)
      t[i] = __ldg(P.via_ids + i);
    for (uint32_t i = threadIdx.x; i < P.via_words; i += blockDim.x)
      bits[i] = __ldg(P.via_bits + i);
    for (uint32_t i = threadIdx.x; i < P.via_pair_bytes / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(pair)[i] = __ldg(reinterpret_cast<const uint32_t*>(P.via_pair) + i);
  }
  return v;
}

// viable positions of an interior chunk (all 19 bytes the test reads exist): bit k = position k may start a match
__device__ __forceinline__ uint32_t viable16(const ViaTables& v, const Window& W)
{
  uint2 a = v.t[UGX_WB(W, 0)], b = v.t[UGX_WB(W, 1)], c = v.t[UGX_WB(W, 2)];
  uint32_t m = 0;
  if (v.bytes)
  {
    const uint8_t* flag = reinterpret_cast<const uint8_t*>(v.bits);
#pragma unroll
    for (int k = 0; k < 16; ++k)
    {
      const uint2 d = v.t[UGX_WB(W, k + 3)];
      const uint32_t code = v.pair[(a.x & 0xffffu) + (b.x >> 16)];
      m += static_cast<uint32_t>(flag[code * v.stride + (c.y & 0xffffu) + (d.y >> 16)]) << k;
      a = b;
      b = c;
      c = d;
    }
    return m;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k)
  {
    const uint2 d = v.t[UGX_WB(W, k + 3)];
    const uint32_t code = v.pair[(a.x & 0xffffu) + (b.x >> 16)];
    const uint32_t idx = code * v.stride + (c.y & 0xffffu) + (d.y >> 16);
    m |= ((v.bits[idx >> 5] >> (idx & 31u)) & 1u) << k;
    a = b;
    b = c;
    c = d;
  }
  return m;
}

} // namespace ugx
