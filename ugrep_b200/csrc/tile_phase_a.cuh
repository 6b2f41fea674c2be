// tile_phase_a.cuh — phase A of every scan kernel: the position-parallel prefilter.
//
// A CTA streams a tile from HBM with coalesced 16-byte loads; each thread holds a 16-byte
// chunk plus an 8-byte halo in registers and evaluates the reference's candidate predicate
// for its 16 positions without touching memory again, except for the predictor tables staged
// in shared memory.  Output per chunk: a 16-bit newline mask and a 16-bit candidate mask,
// stored as bitmaps in shared memory (bit i of the tile bitmap = byte i of the tile).
//
// The register-window ("fast") path is exact wherever every byte it reads exists; chunks
// within 24 bytes of the end of the buffer use the per-position predicate cand(), which
// carries the reference's end-of-buffer rules.
#pragma once

#include "device_pattern.cuh"

namespace ugx {

struct Window {
  uint32_t w[7]; // 24 text bytes (little-endian words) + one zero word of padding
};

// gather the top bit of each byte of a __vcmpeq4 result into a nibble (bit i = byte i)
__device__ __forceinline__ uint32_t nib_of(uint32_t eq) { return ((eq & 0x80808080u) * 0x00204081u) >> 28; }

__device__ __forceinline__ uint32_t newline_mask16(const Window& W)
{
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    m |= nib_of(__vcmpeq4(W.w[i], 0x0a0a0a0au)) << (4 * i);
  return m;
}

// bit k set iff window byte (k + off) is one of chr[first .. first + cnt); off <= 7
__device__ __forceinline__ uint32_t anchor_mask16(const Window& W, const DevPattern& P, uint32_t off, uint32_t first, uint32_t cnt)
{
  const uint32_t q = off >> 2, sh = (off & 3) * 8;
  uint32_t a[4], e[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    const uint32_t lo = q ? W.w[i + 1] : W.w[i];
    const uint32_t hi = q ? W.w[i + 2] : W.w[i + 1];
    a[i] = __funnelshift_r(lo, hi, sh);
    e[i] = 0;
  }
  for (uint32_t n = 0; n < cnt; ++n)
  {
    const uint32_t splat = static_cast<uint32_t>(P.chr[first + n]) * 0x01010101u;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      e[i] |= __vcmpeq4(a[i], splat);
  }
  return nib_of(e[0]) | (nib_of(e[1]) << 4) | (nib_of(e[2]) << 8) | (nib_of(e[3]) << 12);
}

// bytes k..k+3 -> x, k+4..k+7 -> y, k+8..k+11 -> z of the window, k in 0..16
__device__ __forceinline__ void window_at(const Window& W, uint32_t k, uint32_t& x, uint32_t& y, uint32_t& z)
{
  const uint32_t sh = (k & 3) * 8;
  uint32_t a0, a1, a2, a3;
  switch (k >> 2)
  {
    case 0: a0 = W.w[0]; a1 = W.w[1]; a2 = W.w[2]; a3 = W.w[3]; break;
    case 1: a0 = W.w[1]; a1 = W.w[2]; a2 = W.w[3]; a3 = W.w[4]; break;
    case 2: a0 = W.w[2]; a1 = W.w[3]; a2 = W.w[4]; a3 = W.w[5]; break;
    case 3: a0 = W.w[3]; a1 = W.w[4]; a2 = W.w[5]; a3 = W.w[6]; break;
    default: a0 = W.w[4]; a1 = W.w[5]; a2 = W.w[6]; a3 = 0; break;
  }
  x = __funnelshift_r(a0, a1, sh);
  y = __funnelshift_r(a1, a2, sh);
  z = __funnelshift_r(a2, a3, sh);
}

// Pattern::predict_match PM4 on 4 bytes held in x (include/reflex/pattern.h:389-401)
__device__ __forceinline__ bool pm4_x(const uint8_t* __restrict__ pma, uint32_t x)
{
  const uint32_t c0 = x & 0xff, c1 = (x >> 8) & 0xff, c2 = (x >> 16) & 0xff, c3 = x >> 24;
  const uint32_t h1 = hash3(c0, c1), h2 = hash3(h1, c2), h3 = hash3(h2, c3);
  const uint32_t q = (pma[c0] & 0xc0u) | (pma[h1] & 0x30u) | (pma[h2] & 0x0cu) | (pma[h3] & 0x03u);
  const uint32_t r = ((((((q >> 2) | q) >> 2) | q) >> 1) | q) & 0xffu;
  return r != 0xffu;
}

// Pattern::predict_match PMH on n <= 8 bytes held in (x, y) (include/reflex/pattern.h:366-387)
__device__ __forceinline__ bool pmh_xy(const uint8_t* __restrict__ tab, uint32_t x, uint32_t y, uint32_t n)
{
  uint32_t h = x & 0xff;
  if (tab[h] & 1u)
    return false;
  uint64_t v = (static_cast<uint64_t>(y) << 32 | x) >> 8;
  uint32_t bit = 2;
  for (uint32_t j = 1; j < n; ++j, bit <<= 1, v >>= 8)
  {
    h = hash3(h, static_cast<uint32_t>(v) & 0xff);
    if (tab[h] & bit)
      return false;
  }
  return true;
}

#define UGX_WB(W, i) (((W).w[(i) >> 2] >> (((i) & 3) * 8)) & 0xffu)

// the generic per-position path (chunks near the end of the buffer, and exotic routines)
__device__ __forceinline__ uint32_t chunk_cand_generic(const Text& t, const DevPattern& P, const Tables& T, uint64_t base)
{
  uint32_t m = 0;
  for (uint32_t k = 0; k < 16; ++k)
    if (base + k < t.end && cand(t, P, T, base + k))
      m |= 1u << k;
  return m;
}

// candidate mask of an interior chunk (base + 24 <= end) from its register window
__device__ __forceinline__ uint32_t chunk_cand_fast(const Window& W, const Text& t, const DevPattern& P, const Tables& T, uint64_t base)
{
  uint32_t surv = 0;   // positions that still need the predictor
  uint32_t shift = 0;  // predictor looks at position k + shift
  bool use_pmh = false;
  switch (P.adv)
  {
    case UGX_ADV_PIN1_ONE:
      surv = anchor_mask16(W, P, 0, 0, 1);
      break;
    case UGX_ADV_PIN1_PMH:
      use_pmh = true;
      // fall through
    case UGX_ADV_PIN1_PMA:
      surv = anchor_mask16(W, P, P.lcp, 0, 1) & anchor_mask16(W, P, P.lcs, 1, 1);
      break;
    case UGX_ADV_PIN_ONE:
      surv = anchor_mask16(W, P, 0, 0, P.pin);
      break;
    case UGX_ADV_PIN_PMH:
      use_pmh = true;
      // fall through
    case UGX_ADV_PIN_PMA:
      surv = anchor_mask16(W, P, P.lcp, 0, P.pin) & anchor_mask16(W, P, P.lcs, P.pin, P.pin);
      break;
    case UGX_ADV_CHAR:
      return anchor_mask16(W, P, 0, 0, 1);
    case UGX_ADV_CHAR_PMH:
      use_pmh = true;
      // fall through
    case UGX_ADV_CHAR_PMA:
      surv = anchor_mask16(W, P, 0, 0, 1);
      shift = 1;
      break;
    case UGX_ADV_STRING:
    case UGX_ADV_STRING_PMA:
    case UGX_ADV_STRING_PMH:
    {
      // two anchor bytes of the literal inside the window, then the exact test on the few survivors
      const uint32_t a0 = P.lcp < 8 ? P.lcp : 0;
      const uint32_t a1 = P.lcs < 8 ? P.lcs : (P.len - 1 < 7 ? P.len - 1 : 7);
      uint32_t s = anchor_mask16(W, P, a0, a0, 1) & anchor_mask16(W, P, a1, a1, 1);
      uint32_t m = 0;
      while (s != 0)
      {
        const uint32_t k = __ffs(s) - 1;
        s &= s - 1;
        if (cand(t, P, T, base + k))
          m |= 1u << k;
      }
      return m;
    }
    case UGX_ADV_MIN1:
    case UGX_ADV_MIN2:
    case UGX_ADV_MIN3:
    case UGX_ADV_MIN4:
    {
      // bitap over hashed byte pairs: position k survives if pair (k+j, k+j+1) has bit j clear for all j < min
      const uint32_t depth = P.min < 1 ? 1 : P.min;
      uint32_t bad = 0;
#pragma unroll
      for (int i = 0; i < 23; ++i)
      {
        if (i < 15 + static_cast<int>(depth))
        {
          const uint32_t tp = T.tap[bihash(UGX_WB(W, i), UGX_WB(W, i + 1))];
          // pair i is step j = i - k of position k
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < static_cast<int>(depth) && i - j >= 0 && i - j < 16)
              bad |= ((tp >> j) & 1u) << (i - j);
        }
      }
      surv = ~bad & 0xffffu;
      use_pmh = P.adv == UGX_ADV_MIN4;
      break;
    }
    case UGX_ADV_PMA:
    {
      uint32_t m = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k)
      {
        const uint32_t c0 = UGX_WB(W, k), c1 = UGX_WB(W, k + 1), c2 = UGX_WB(W, k + 2), c3 = UGX_WB(W, k + 3);
        const uint32_t h1 = hash3(c0, c1), h2 = hash3(h1, c2), h3 = hash3(h2, c3);
        const uint32_t q = (T.pred[c0] & 0xc0u) | (T.pred[h1] & 0x30u) | (T.pred[h2] & 0x0cu) | (T.pred[h3] & 0x03u);
        const uint32_t r = ((((((q >> 2) | q) >> 2) | q) >> 1) | q) & 0xffu;
        m |= (r != 0xffu ? 1u : 0u) << k;
      }
      return m;
    }
    default:
      return chunk_cand_generic(t, P, T, base);
  }
  uint32_t m = 0;
  while (surv != 0)
  {
    const uint32_t k = __ffs(surv) - 1;
    surv &= surv - 1;
    uint32_t x, y, z;
    window_at(W, k + shift, x, y, z);
    const bool ok = use_pmh ? pmh_xy(T.pred, x, y, P.min) : pm4_x(T.pred, x);
    if (ok)
      m |= 1u << k;
  }
  return m;
}

// Load chunk `base` (16-byte aligned, base < end) and its halo; returns false if the chunk is not interior
__device__ __forceinline__ bool load_window(const uint8_t* __restrict__ buf, uint64_t end, uint64_t base, Window& W)
{
  W.w[6] = 0;
  if (base + 24 <= end)
  {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + base));
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(buf + base + 16));
    W.w[0] = v.x; W.w[1] = v.y; W.w[2] = v.z; W.w[3] = v.w; W.w[4] = h.x; W.w[5] = h.y;
    return true;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
  {
    uint32_t x = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
    {
      const uint64_t p = base + 4 * i + b;
      if (p < end)
        x |= static_cast<uint32_t>(__ldg(buf + p)) << (8 * b);
    }
    W.w[i] = x;
  }
  return false;
}

// Phase A for one tile: CHUNKS chunks per thread, strided by the block size so that a warp's loads coalesce.
// Writes the tile's candidate and newline bitmaps (16 bits per chunk) and returns this thread's newline count.
template <int CHUNKS>
__device__ __forceinline__ uint32_t tile_phase_a(const Text& t, const DevPattern& P, const Tables& T, uint64_t tile_base,
                                                 uint16_t* __restrict__ cand16, uint16_t* __restrict__ nl16)
{
  uint32_t nlcount = 0;
  const uint32_t lane = threadIdx.x & 31;
  // issue all loads of the tile first (memory-level parallelism), then evaluate
  uint4 v[CHUNKS];
#pragma unroll
  for (int j = 0; j < CHUNKS; ++j)
  {
    const uint64_t base = tile_base + (static_cast<uint64_t>(j) * blockDim.x + threadIdx.x) * 16;
    if (base + 16 <= t.end)
      v[j] = __ldg(reinterpret_cast<const uint4*>(t.b + base));
    else
    {
      // the last, partial chunk of the buffer: assemble it bytewise (the next lane's halo comes from here)
      uint32_t x[4] = {0, 0, 0, 0};
      for (uint32_t i = 0; i < 16 && base + i < t.end; ++i)
        x[i >> 2] |= static_cast<uint32_t>(__ldg(t.b + base + i)) << (8 * (i & 3));
      v[j] = make_uint4(x[0], x[1], x[2], x[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < CHUNKS; ++j)
  {
    const uint32_t g = j * blockDim.x + threadIdx.x;
    const uint64_t base = tile_base + static_cast<uint64_t>(g) * 16;
    // halo: the first 8 bytes of the next chunk live in the next lane's registers
    uint32_t hx = __shfl_down_sync(0xffffffffu, v[j].x, 1);
    uint32_t hy = __shfl_down_sync(0xffffffffu, v[j].y, 1);
    uint32_t cm = 0, nl = 0;
    if (base < t.end)
    {
      Window W;
      bool interior = base + 24 <= t.end;
      if (interior)
      {
        if (lane == 31)
        {
          const uint2 h = __ldg(reinterpret_cast<const uint2*>(t.b + base + 16));
          hx = h.x;
          hy = h.y;
        }
        W.w[0] = v[j].x; W.w[1] = v[j].y; W.w[2] = v[j].z; W.w[3] = v[j].w; W.w[4] = hx; W.w[5] = hy; W.w[6] = 0;
        cm = chunk_cand_fast(W, t, P, T, base);
      }
      else
      {
        load_window(t.b, t.end, base, W);
        cm = chunk_cand_generic(t, P, T, base);
      }
      nl = newline_mask16(W);
      if (base + 16 > t.end)
        nl &= (1u << (t.end - base)) - 1;
    }
    cand16[g] = static_cast<uint16_t>(cm);
    nl16[g] = static_cast<uint16_t>(nl);
    nlcount += __popc(nl);
  }
  return nlcount;
}

} // namespace ugx
