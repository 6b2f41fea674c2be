// stream_common.cuh — pieces shared by the streaming count kernels (stream_literal.cu, stream_count.cu):
// SWAR byte tests, the per-warp line state with its carry-arithmetic resolution, the region ticket and
// the CTA epilogue in which the last CTA chains the regions' head lines.  See stream_count.cu for the
// decomposition.
#pragma once

#include "device_pattern.cuh"
#include "scan_kernels.hpp"

namespace ugx {

constexpr uint32_t SC_SPAN = 512; // bytes per span: 32 lanes x 16

// ---- SWAR byte tests --------------------------------------------------------------------------------
// exact: 0x80 in every byte of the result whose byte in x is zero
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x)
{
  const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return ~(t | x) & 0x80808080u;
}

// accumulating "some byte of x is zero" test: the 0x80 bits of the result are exact up to and including
// the lowest zero byte (bits above it may be false positives), so `!= 0` after masking is exact
__device__ __forceinline__ uint32_t zero_any(uint32_t x, uint32_t acc) { return ((x - 0x01010101u) & ~x) | acc; }

// gather the 0x80 flags of four bytes into a nibble (bit i = byte i)
__device__ __forceinline__ uint32_t flags_to_nibble(uint32_t f) { return (f * 0x00204081u) >> 28; }

__device__ __forceinline__ uint32_t load_bytes(const uint8_t* __restrict__ buf, uint64_t n, uint64_t at)
{
  uint32_t x = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (at + b < n)
      x |= static_cast<uint32_t>(__ldg(buf + at + b)) << (8 * b);
  return x;
}

// a 16-byte chunk that may straddle or lie past the end of the buffer: bytes at or past n read as zero
__device__ __forceinline__ uint4 load_chunk_guarded(const uint8_t* __restrict__ buf, uint64_t n, uint64_t base)
{
  if (base + 16 <= n)
    return __ldg(reinterpret_cast<const uint4*>(buf + base));
  return make_uint4(load_bytes(buf, n, base), load_bytes(buf, n, base + 4), load_bytes(buf, n, base + 8),
                    load_bytes(buf, n, base + 12));
}

__device__ __forceinline__ uint32_t newline_mask_exact4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
  return flags_to_nibble(zero_bytes(w0 ^ 0x0a0a0a0au)) | (flags_to_nibble(zero_bytes(w1 ^ 0x0a0a0a0au)) << 4) |
         (flags_to_nibble(zero_bytes(w2 ^ 0x0a0a0a0au)) << 8) | (flags_to_nibble(zero_bytes(w3 ^ 0x0a0a0a0au)) << 12);
}

// ---- per-warp line state -----------------------------------------------------------------------------
struct LineState {
  uint32_t cin;      // the line open at the cursor already has a success (starts at 1: the region's head line is deferred)
  bool seen_nl;      // a newline was seen in this region
  bool watch;        // newlines matter right now: cin is set or the region's first newline is still to come
  bool head;         // success before the first newline of the region
  uint32_t ucount;   // warp-uniform count (added once per warp)
  uint32_t lcount;   // lane-private count
};

// a span with at least one success: exact masks per lane, carry chain across lanes
__device__ __forceinline__ void resolve_span(LineState& L, uint32_t nl16, uint32_t succ16)
{
  const uint32_t s = succ16 & ~nl16;
  const uint32_t v = (s + (~nl16 & 0xffffu)) & (nl16 | 0x10000u);
  const uint32_t first = nl16 & (0u - nl16);
  const bool has = nl16 != 0;
  const uint32_t g = v >> 16;                                // success after the last newline (or anywhere, if none)
  const bool hs = has ? (v & first) != 0 : g != 0;           // success before the first newline
  L.lcount += __popc(v & nl16 & ~first) + (has ? g : 0u);    // lines that start inside this chunk
  const uint32_t NL = __ballot_sync(0xffffffffu, has);
  const uint32_t H = __ballot_sync(0xffffffffu, hs);
  const uint32_t G = __ballot_sync(0xffffffffu, g != 0);
  const uint32_t A = G | ~NL;
  const uint64_t sum = static_cast<uint64_t>(A) + G + L.cin;
  const uint32_t C = static_cast<uint32_t>(sum) ^ A ^ G;     // bit l: the line open at lane l's first byte already counted
  if (!L.seen_nl)
  {
    const uint32_t headlanes = NL != 0 ? (((NL & (0u - NL)) << 1) - 1u) : 0xffffffffu;
    if ((H & headlanes) != 0)
      L.head = true;
  }
  L.ucount += __popc(H & ~C);
  L.cin = static_cast<uint32_t>(sum >> 32);
  if (NL != 0)
    L.seen_nl = true;
  L.watch = L.cin != 0 || !L.seen_nl;
}

// Region schedule of one warp: the first `rounds` regions are static (round i: region i * total_warps + warp),
// the remaining eighth of the buffer is handed out by an atomic ticket so that the tail balances.  A single
// address takes only ~0.3 G atomics/s on B200 — one ticket per 16 KiB region for the whole buffer would cap the
// kernel at ~4.7 TB/s (tools/probe/read_bw.cu) — hence the static part.
struct RegionSchedule {
  unsigned long long* ticket;
  uint64_t total_warps, warp, rounds, static_regions;
  uint64_t i;
  __device__ __forceinline__ void init(unsigned long long* t, uint64_t nregions)
  {
    ticket = t;
    total_warps = static_cast<uint64_t>(gridDim.x) * (blockDim.x >> 5);
    warp = static_cast<uint64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    rounds = (nregions - nregions / 8) / total_warps;
    static_regions = rounds * total_warps;
    i = 0;
  }
  // warp-uniform; >= nregions when the buffer is exhausted
  __device__ __forceinline__ uint64_t next(uint32_t lane)
  {
    if (i < rounds)
      return (i++) * total_warps + warp;
    unsigned long long tk = 0;
    if (lane == 0)
      tk = atomicAdd(ticket, 1ull);
    return static_regions + __shfl_sync(0xffffffffu, tk, 0);
  }
};

#ifndef UGX_SC_PF
#define UGX_SC_PF 0
#endif
// ask the TMA engine to pull `bytes` at p into L2 (no destination: a prefetch), one thread per request
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// CTA epilogue: publish the CTA's partial counts; the last CTA to arrive sums the partials and, on the final
// launch of a buffer, chains the region summaries (bit 0 has newline, bit 1 head success, bit 2 carry out).
__device__ __forceinline__ void stream_epilogue(const StreamArgs& a, uint64_t nregions, unsigned long long my_lines,
                                                unsigned long long my_newlines, uint32_t warp_uniform_lines)
{
  __shared__ unsigned long long s_red[2 * 32];
  __shared__ uint32_t s_last;
  __shared__ uint8_t s_slice[1024];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
    my_lines += warp_uniform_lines;

  // ---- CTA partials
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
  {
    my_lines += __shfl_down_sync(0xffffffffu, my_lines, d);
    my_newlines += __shfl_down_sync(0xffffffffu, my_newlines, d);
  }
  if (lane == 0)
  {
    s_red[2 * wid] = my_lines;
    s_red[2 * wid + 1] = my_newlines;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned long long x = 0, y = 0;
    for (uint32_t i = 0; i < blockDim.x / 32; ++i)
    {
      x += s_red[2 * i];
      y += s_red[2 * i + 1];
    }
    a.partials[2 * blockIdx.x] = x;
    a.partials[2 * blockIdx.x + 1] = y;
    __threadfence();
    const unsigned int done = atomicAdd(a.done, 1u);
    s_last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last)
    return;
  // ---- the last CTA: add the partials, and on the final launch of a buffer chain the regions' head lines
  __threadfence();
  unsigned long long lines = 0, newlines = 0;
  for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x)
  {
    lines += __ldcg(a.partials + 2 * i);
    newlines += __ldcg(a.partials + 2 * i + 1);
  }
  long long adjust = 0;
  if (a.finalize)
  {
    // thread i chains the slice [lo, hi) assuming no carry in; the slices are then chained serially
    const uint64_t total = nregions;
    const uint64_t per = ((total + blockDim.x - 1) / blockDim.x + 15) & ~15ull;
    const uint64_t lo = threadIdx.x * per;
    const uint64_t hi = lo + per < total ? lo + per : total;
    uint32_t c = 0, seen = 0, hbf = 0;
    for (uint64_t i = lo; i < hi; i += 16)
    {
      const uint4 q = __ldcg(reinterpret_cast<const uint4*>(a.region_sum + i));
      const uint32_t ww[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 16; ++k)
      {
        if (i + k < hi)
        {
          const uint32_t bsum = (ww[k >> 2] >> (8 * (k & 3))) & 0xffu;
          const uint32_t nl = bsum & 1u, hd = (bsum >> 1) & 1u, g = (bsum >> 2) & 1u;
          if (hd && !c)
          {
            ++adjust;
            if (!seen)
              hbf = 1;
          }
          c = nl ? g : (c | hd);
          seen |= nl;
        }
      }
    }
    s_slice[threadIdx.x] = static_cast<uint8_t>(c | (seen << 1) | (hbf << 2));
  }
  __syncthreads();
  if (a.finalize && threadIdx.x == 0)
  {
    uint32_t cin = 0;
    for (uint32_t i = 0; i < blockDim.x; ++i)
    {
      const uint32_t sl = s_slice[i];
      if (cin && (sl & 4u))
        --adjust;
      cin = (sl & 2u) ? (sl & 1u) : (cin | (sl & 1u));
    }
  }
  lines += static_cast<unsigned long long>(adjust);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
  {
    lines += __shfl_down_sync(0xffffffffu, lines, d);
    newlines += __shfl_down_sync(0xffffffffu, newlines, d);
  }
  __syncthreads();
  if (lane == 0)
  {
    s_red[2 * wid] = lines;
    s_red[2 * wid + 1] = newlines;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    unsigned long long x = 0, y = 0;
    for (uint32_t i = 0; i < blockDim.x / 32; ++i)
    {
      x += s_red[2 * i];
      y += s_red[2 * i + 1];
    }
    if (a.accumulate)
    {
      x += a.totals[0];
      y += a.totals[1];
    }
    a.totals[0] = x;
    a.totals[1] = y;
    *a.done = 0;
    *a.ticket = 0;
  }
}


// ---- the streaming skeleton ---------------------------------------------------------------------------
// Warps take 16 KiB regions from the ticket counter and walk them in 2 KiB blocks of four 512-byte spans.
// Lane l owns chunk l (16 bytes) of a span.  The four chunk registers v[0..3] form a ring: as soon as span j
// of the current block has been evaluated, v[j] is reloaded with span j of the NEXT block (the next block of
// the region, or the first block of the warp's next region), so about 2 KiB per warp are always in flight and
// only 16 data registers are live.  The 12-byte halo a chunk needs from its right neighbour comes from the
// next lane by shuffle; lane 31 takes it from lane 0 of the next span, or from `h`, the 16 bytes after the block.
//
// Eval: bool operator()(const uint32_t (&w)[7], uint64_t sbase, uint32_t& succ16) — evaluates one span (w = the
// lane's 16 bytes + 12 halo bytes), sets the lane's 16-bit success mask and returns the
// warp-uniform "some lane has a success".

// one span.  WATCH: a newline would change the line state (cin is set, or the region's first newline is still
// to come), or newlines are being counted; otherwise the span is only tested for successes.
template <bool WATCH, bool WANT_NL, class Eval>
__device__ __forceinline__ void stream_span(const uint32_t (&w)[7], uint64_t sbase, LineState& L, uint32_t& nlacc, Eval& ev)
{
  uint32_t nl_any = 0;
  if (WANT_NL)
  {
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const uint32_t e = zero_bytes(w[i] ^ 0x0a0a0a0au);
      nlacc += e >> 7;
      nl_any |= e;
    }
  }
  else if (WATCH)
  {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      nl_any = zero_any(w[i] ^ 0x0a0a0a0au, nl_any);
    nl_any &= 0x80808080u;
  }
  uint32_t succ16 = 0;
  if (ev(w, sbase, succ16))
  {
    resolve_span(L, newline_mask_exact4(w[0], w[1], w[2], w[3]), succ16);
  }
  else if (WATCH && L.watch && __any_sync(0xffffffffu, nl_any != 0))
  {
    L.seen_nl = true;
    L.cin = 0;
    L.watch = false;
  }
}

// one region.  FULL: the region, its halo and the first block of the warp's next region lie wholly inside the
// buffer, so no load is guarded and no span is tested against the end of the buffer.
// SC_SPANS = spans per block (4 for the memory-bound literal kernel; 1 for the DFA kernels, whose span evaluation
// is large and must not be replicated four times in the instruction stream).
template <bool FULL, bool WANT_NL, bool SPLIT_WATCH, int SC_SPANS, class Eval>
__device__ __forceinline__ void stream_region(const uint8_t* __restrict__ buf, uint64_t n, uint64_t rbase, bool have_next,
                                              uint64_t next_rbase, uint32_t lane, uint32_t next_lane, uint4 (&v)[SC_SPANS],
                                              uint4& h, LineState& L, unsigned long long& my_newlines, Eval& ev)
{
  constexpr uint32_t SC_BLOCK = SC_SPANS * SC_SPAN;
  const uint64_t rend = FULL ? rbase + SC_REGION : (rbase + SC_REGION < n ? rbase + SC_REGION : n);
  const uint32_t nblocks = FULL ? SC_REGION / SC_BLOCK : static_cast<uint32_t>((rend - rbase + SC_BLOCK - 1) / SC_BLOCK);
  const uint8_t* __restrict__ p = buf + rbase + lane * 16; // this lane's chunk 0 of the current block
  const uint8_t* pnr = buf + next_rbase + lane * 16;        // ... and of the first block of the warp's next region
  asm volatile("" : "+l"(pnr));                              // (kept in registers: not recomputed per block)
  uint32_t nlacc = 0;                                      // WANT_NL: per-byte-lane newline counters, flushed per block
  for (uint32_t b = 0; b < nblocks; ++b)
  {
    const uint64_t bbase = rbase + static_cast<uint64_t>(b) * SC_BLOCK;
    const bool last = b + 1 == nblocks;
    const bool has_next = !last || have_next;
    const uint64_t nbase = last ? next_rbase : bbase + SC_BLOCK;
    const uint8_t* pn = last ? pnr : p + SC_BLOCK;
    asm volatile("" : "+l"(pn)); // keep the reload pointer in registers: ptxas otherwise recomputes it for every span
    const bool next_full = FULL || nbase + SC_BLOCK + 16 <= n;
    if (FULL && UGX_SC_PF > 0 && lane == 0)
    {
      // L2 prefetch UGX_SC_PF blocks ahead: within the region, then into the warp's next region
      const uint32_t ahead = b + UGX_SC_PF;
      if (ahead < nblocks)
        l2_prefetch(buf + rbase + static_cast<uint64_t>(ahead) * SC_BLOCK, SC_BLOCK);
      else if (have_next && next_rbase + static_cast<uint64_t>(ahead - nblocks + 1) * SC_BLOCK <= n)
        l2_prefetch(buf + next_rbase + static_cast<uint64_t>(ahead - nblocks) * SC_BLOCK, SC_BLOCK);
    }
    uint4 hn = h;
    if (has_next)
    {
      if (next_full)
        hn = __ldg(reinterpret_cast<const uint4*>(pn - lane * 16 + SC_BLOCK));
      else
        hn = load_chunk_guarded(buf, n, nbase + SC_BLOCK);
    }
#pragma unroll
    for (int j = 0; j < SC_SPANS; ++j)
    {
      const uint64_t sbase = bbase + j * SC_SPAN;
      if (FULL || sbase < n)
      {
        uint32_t w[7];
        w[0] = v[j].x;
        w[1] = v[j].y;
        w[2] = v[j].z;
        w[3] = v[j].w;
        const uint4 nv = j + 1 < SC_SPANS ? v[j + 1 < SC_SPANS ? j + 1 : j] : h;
        w[4] = __shfl_sync(0xffffffffu, lane == 0 ? nv.x : w[0], next_lane);
        w[5] = __shfl_sync(0xffffffffu, lane == 0 ? nv.y : w[1], next_lane);
        w[6] = __shfl_sync(0xffffffffu, lane == 0 ? nv.z : w[2], next_lane);
        // two instantiations so that the cruising path (no newline watch) carries no newline code at all
        // (SPLIT_WATCH false: kernels whose span evaluation dwarfs the newline test keep one copy, for code size)
        if (WANT_NL || !SPLIT_WATCH || L.watch)
          stream_span<true, WANT_NL>(w, sbase, L, nlacc, ev);
        else
          stream_span<false, WANT_NL>(w, sbase, L, nlacc, ev);
      }
      if (has_next)
        v[j] = next_full ? __ldg(reinterpret_cast<const uint4*>(pn + j * SC_SPAN))
                         : load_chunk_guarded(buf, n, nbase + j * SC_SPAN + lane * 16);
    }
    h = hn;
    p += SC_BLOCK;
    if (WANT_NL)
    {
      const uint32_t pair = (nlacc & 0x00ff00ffu) + ((nlacc >> 8) & 0x00ff00ffu);
      my_newlines += (pair & 0xffffu) + (pair >> 16);
      nlacc = 0;
    }
  }
}

template <bool WANT_NL, bool SPLIT_WATCH, int SC_SPANS, class Eval>
__device__ __forceinline__ void stream_scan(const uint8_t* __restrict__ buf, uint64_t n, const StreamArgs& a, Eval& ev)
{
  constexpr uint32_t SC_BLOCK = SC_SPANS * SC_SPAN;
  uint32_t lane = threadIdx.x & 31;
  uint32_t next_lane = (lane + 1) & 31;
  asm volatile("" : "+r"(lane), "+r"(next_lane)); // opaque: no re-reading of %tid in the hot loop
  // this launch scans the regions [region_begin, region_end) of the buffer; n = bytes of the buffer that are valid
  // (a host buffer is scanned while it is still being copied: capi.cu)
  const uint64_t nregions = a.region_end;
  unsigned long long my_lines = 0, my_newlines = 0;
  uint32_t warp_uniform_lines = 0;
  RegionSchedule sched;
  sched.init(a.ticket, a.region_end - a.region_begin);
  uint64_t r = a.region_begin + sched.next(lane);
  uint4 v[SC_SPANS];
  uint4 h = make_uint4(0, 0, 0, 0);
  if (r < nregions)
  {
    const uint64_t b0 = r * SC_REGION;
#pragma unroll
    for (int j = 0; j < SC_SPANS; ++j)
      v[j] = load_chunk_guarded(buf, n, b0 + j * SC_SPAN + lane * 16);
    h = load_chunk_guarded(buf, n, b0 + SC_BLOCK);
  }
  while (r < nregions)
  {
    const uint64_t rbase = r * SC_REGION;
    if (rbase >= n)
      break; // defensive: a region range that does not match the valid bytes must not spin
    const uint64_t rnext = a.region_begin + sched.next(lane);
    const bool have_next = rnext < nregions;
    const uint64_t next_rbase = rnext * SC_REGION;
    LineState L;
    L.cin = 1;
    L.seen_nl = false;
    L.watch = true;
    L.head = false;
    L.ucount = 0;
    L.lcount = 0;
    const bool full = rbase + SC_REGION + 16 <= n && (!have_next || next_rbase + SC_BLOCK + 16 <= n);
    if (full)
      stream_region<true, WANT_NL, SPLIT_WATCH, SC_SPANS>(buf, n, rbase, have_next, next_rbase, lane, next_lane, v, h, L, my_newlines, ev);
    else
      stream_region<false, WANT_NL, SPLIT_WATCH, SC_SPANS>(buf, n, rbase, have_next, next_rbase, lane, next_lane, v, h, L, my_newlines, ev);
    // publish the region: bit 0 has newline, bit 1 head success, bit 2 carry out
    if (lane == 0)
    {
      const uint32_t g = L.seen_nl ? L.cin : (L.head ? 1u : 0u);
      a.region_sum[r] = static_cast<uint8_t>((L.seen_nl ? 1u : 0u) | (L.head ? 2u : 0u) | (g << 2));
    }
    my_lines += L.lcount;
    warp_uniform_lines += L.ucount;
    r = rnext;
  }
  stream_epilogue(a, nregions, my_lines, my_newlines, warp_uniform_lines);
}

} // namespace ugx
