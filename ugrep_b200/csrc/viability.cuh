// viability.cuh — the k-gram viability test of an anchored attempt (pattern_host.hpp Viability): a position whose
// next k <= 4 bytes cannot keep the DFA alive (nor make it accept) starts no match, so the scan kernels spare its
// attempt.  Two shared-memory lookups per position; a property of the DFA alone, it never changes a result.
#pragma once

#include "device_pattern.cuh"
#include "tile_phase_a.cuh"

namespace ugx {

struct ViaTables {
  const uint32_t* ids;  // [256] shared
  const uint8_t* pair;  // shared
  const uint32_t* bits; // shared
  uint32_t n1, n2, n3;
  bool on;
};

__host__ __device__ inline uint32_t via_smem_bytes(const DevPattern& P) { return P.via_k ? 1024 + P.via_pair_bytes + P.via_words * 4 : 0; }

// stage the tables behind `base` (16-byte aligned); every thread of the CTA calls it; the caller synchronises
__device__ __forceinline__ ViaTables via_stage(const DevPattern& P, uint8_t* base, bool on)
{
  ViaTables v;
  v.on = on && P.via_k != 0;
  v.n1 = P.via_n[1];
  v.n2 = P.via_n[2];
  v.n3 = P.via_n[3];
  uint32_t* ids = reinterpret_cast<uint32_t*>(base);
  uint32_t* bits = ids + 256;
  uint8_t* pair = reinterpret_cast<uint8_t*>(bits + P.via_words);
  v.ids = ids;
  v.bits = bits;
  v.pair = pair;
  if (v.on)
  {
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x)
      ids[i] = __ldg(P.via_ids + i);
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
  uint32_t i0 = v.ids[UGX_WB(W, 0)], i1 = v.ids[UGX_WB(W, 1)], i2 = v.ids[UGX_WB(W, 2)];
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k)
  {
    const uint32_t i3 = v.ids[UGX_WB(W, k + 3)];
    const uint32_t code = v.pair[(i0 & 0xffu) * v.n1 + ((i1 >> 8) & 0xffu)];
    const uint32_t c = code == 255u ? 0u : code;
    const uint32_t idx = (c * v.n2 + ((i2 >> 16) & 0xffu)) * v.n3 + (i3 >> 24);
    const uint32_t bit = (v.bits[idx >> 5] >> (idx & 31u)) & 1u;
    m |= (code == 255u ? 1u : (code != 0u ? bit : 0u)) << k;
    i0 = i1;
    i1 = i2;
    i2 = i3;
  }
  return m;
}

} // namespace ugx
