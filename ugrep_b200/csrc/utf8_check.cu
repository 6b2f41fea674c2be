// utf8_check.cu — utf8_check_kernel: reflex::isutf8 (lib/simd.cpp:169-421, AVX2 form lib/simd_avx2.cpp:82-149), the
// test behind ugrep's binary-file detection (is_binary() = !isutf8(), src/ugrep.cpp:699-711; Grep::init_is_binary,
// :3998-4017), and the NUL test ugrep uses instead with -U (memchr(s, 0, n)).
//
// The reference's rule is position-local once the three bytes before a position are known: with
//   L2 = byte is 11xxxxxx, L3 = 111xxxxx, L4 = 1111xxxx, C = 10xxxxxx        (p, q, r of the reference's scalar form)
// a continuation byte is EXPECTED at j iff L2(j-1) | L3(j-2) | L4(j-3), and the text is valid iff at every j
// "expected" equals C(j) and the byte is none of 00, C0, C1, F5..FF.  A sequence cut off by the end of the buffer shows
// up as an expectation on the first bytes past the end, which are read as 'A'.  (Overlong 3- / 4-byte forms and
// surrogates pass, as they do in the reference.)  One streaming pass, 16 bytes per lane per load, SWAR on the 0x80 bits:
// HBM-bound.
#include <cstdint>
#include <cuda_runtime.h>

#include "scan_kernels.hpp"
#include "stream_common.cuh"

namespace ugx {

namespace {

constexpr uint32_t H = 0x80808080u;

struct ByteMasks {
  uint32_t c, l2, l3, l4; // 0x80 flag per byte
};

__device__ __forceinline__ ByteMasks classify(uint32_t w)
{
  ByteMasks m;
  const uint32_t hi = w & H, b6 = (w << 1) & H, b5 = (w << 2) & H, b4 = (w << 3) & H;
  m.c = hi & ~b6;
  m.l2 = hi & b6;
  m.l3 = m.l2 & b5;
  m.l4 = m.l3 & b4;
  return m;
}

// 0x80 flags of the bytes that may never occur: 00, C0, C1, F5..FF
__device__ __forceinline__ uint32_t forbidden(uint32_t w, const ByteMasks& m, uint32_t& nul)
{
  const uint32_t z = zero_bytes(w);
  nul |= z;
  const uint32_t c0c1 = zero_bytes((w & 0xfefefefeu) ^ 0xc0c0c0c0u);
  const uint32_t b3 = (w << 4) & H, b2 = (w << 5) & H, b1 = (w << 6) & H, b0 = (w << 7) & H;
  return z | c0c1 | (m.l4 & (b3 | (b2 & (b1 | b0))));
}

} // namespace

// flags[0]: bit 0 = not valid UTF-8 by the reference's rule, bit 1 = a NUL byte occurs
__global__ void __launch_bounds__(256) utf8_check_kernel(const uint8_t* __restrict__ buf, uint64_t n, unsigned int* __restrict__ flags)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  // the three positions after the end can still hold an expectation: they are scanned too, as 'A's
  const uint64_t nspans = (n + 3 + 511) / 512;
  uint32_t err = 0, nul = 0;
  for (uint64_t sp = warp; sp < nspans; sp += nwarps)
  {
    const uint64_t base = sp * 512 + lane * 16;
    uint4 v;
    if (base + 16 <= n)
      v = __ldg(reinterpret_cast<const uint4*>(buf + base));
    else
    {
      uint32_t x[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
      for (uint32_t i = 0; i < 16 && base + i < n; ++i)
        x[i >> 2] = (x[i >> 2] & ~(0xffu << (8 * (i & 3)))) | (static_cast<uint32_t>(__ldg(buf + base + i)) << (8 * (i & 3)));
      v = make_uint4(x[0], x[1], x[2], x[3]);
    }
    // the word before my chunk: the previous lane's last word; lane 0 reads it (the buffer starts after an 'A')
    uint32_t prev = __shfl_up_sync(0xffffffffu, v.w, 1);
    if (lane == 0)
      prev = base >= 4 && base - 4 + 4 <= n ? __ldg(reinterpret_cast<const uint32_t*>(buf + base - 4)) : 0x41414141u;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    ByteMasks pm = classify(prev);
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const ByteMasks m = classify(w[i]);
      const uint32_t expect = __funnelshift_l(pm.l2, m.l2, 8) | __funnelshift_l(pm.l3, m.l3, 16) | __funnelshift_l(pm.l4, m.l4, 24);
      err |= (expect ^ m.c) | forbidden(w[i], m, nul);
      pm = m;
    }
  }
  // bytes past the end were read as 'A': they cannot raise the NUL flag, and raise the error flag only through an
  // expectation left by the real text
  const bool e = __any_sync(0xffffffffu, (err & H) != 0), z = __any_sync(0xffffffffu, (nul & H) != 0);
  if (lane == 0 && (e || z))
    atomicOr(flags, (e ? 1u : 0u) | (z ? 2u : 0u));
}

cudaError_t launch_utf8_check(const uint8_t* buf, uint64_t n, unsigned int* flags, int sm_count, cudaStream_t st)
{
  const uint64_t spans = (n + 3 + 511) / 512;
  uint64_t g = (spans + 7) / 8;
  const uint64_t cap = static_cast<uint64_t>(sm_count) * 8;
  if (g > cap)
    g = cap;
  if (g == 0)
    g = 1;
  utf8_check_kernel<<<static_cast<int>(g), 256, 0, st>>>(buf, n, flags);
  return cudaGetLastError();
}

} // namespace ugx
