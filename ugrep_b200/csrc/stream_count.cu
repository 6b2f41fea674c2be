// stream_count.cu — count_lines_stream_kernel: `ugrep -c` (lines with at least one match) for patterns
// without look-back, as ONE streaming pass with no block-wide synchronisation in the hot loop.
//
// Why position-parallel is exact here: without look-back (lbk_ == 0) and without option W, a line
// matches iff SOME prefilter candidate in it starts a non-empty DFA match — the reference's find loop
// (lib/matcher.cpp:42-750) tries the candidates of a line in order until one succeeds and `ugrep -c`
// then skips to the next line (src/ugrep.cpp:10567-10586).  So "success at byte k" is a position-local
// predicate and a line counts iff it holds somewhere in the line.
//
// Work decomposition (DESIGN.md "streaming count"):
//   * the buffer is cut into 16 KiB REGIONS; warps grab regions from an atomic ticket counter;
//   * a warp walks its region in 2 KiB BLOCKS of four 512-byte SPANS: lane l owns the 16-byte chunk l of
//     a span (one coalesced LDG.128 per lane per span), the next block is in flight while the current
//     one is evaluated; the 8-byte halo of a chunk comes from the next lane by shuffle;
//   * per chunk a branch-free SWAR test answers "any newline?" and "any candidate?"; a span without a
//     success costs two ballots.  Only spans with a success build exact 16-bit newline / success masks
//     and resolve "first success of its line" with carry arithmetic: within a lane
//         v = (succ + ~nl) & (nl | bit16)
//     leaves a 1 at every newline whose line segment holds a success (a carry starts at a success and
//     runs up to the next newline), across lanes the same recurrence  c' = g | (p & c)  is one 64-bit
//     add of ballots;
//   * the line that is open at the start of a region cannot be resolved locally: the region publishes
//     three bits (has newline, success before the first newline, success after the last newline) and
//     the last CTA to finish chains them over all regions.
#include "device_pattern.cuh"
#include "line_match.cuh"
#include "scan_kernels.hpp"
#include "ptx.cuh"
#include "stream_common.cuh"
#include "tile_phase_a.cuh"
#include "viability.cuh"

namespace ugx {

namespace {

enum StreamKind { SK_TABLE = 1, SK_META = 2 };

// ---- anchored DFA attempts ---------------------------------------------------------------------------
// does an anchored attempt at `pos` give a non-empty match?  Dense table, states numbered so that
// "accepting" is one compare (pattern_host.hpp): the first accepting state reached decides.
__device__ __forceinline__ bool attempt_table(const Text& t, const DevPattern& P, const Tables& T, uint64_t pos)
{
  uint32_t state = 0;
  uint64_t p = pos;
  for (;;)
  {
    if (p >= t.end)
      return false;
    const uint32_t ch = t.raw(p++);
    const uint32_t nxt = T.next[state * P.ncls + T.cls[ch]];
    if (nxt >= P.first_acc)
    {
      if (nxt == D_DEAD)
        return false;
      if (nxt < P.first_leaf)
        return true;
      return (__ldg(P.accept + nxt) & 0x7fffffffu) != 0; // leaf: accepting unless it is a dead end
    }
    if (nxt == 0 && P.acc0)
      return true;
    state = nxt;
  }
}

__device__ __noinline__ bool attempt_meta(const Text& t, const DevPattern& P, uint64_t pos)
{
  Cursor m;
  set_current(t, m, pos);
  m.txt = pos;
  m.len = 0;
  uint32_t retry = 0;
  run_dfa_opc(t, P, m, retry);
  return m.cap != 0 && m.cur > m.txt;
}

// the same attempt when the first 8 text bytes at `pos` are already in registers (pos + 8 <= end)
__device__ __forceinline__ bool attempt_table_first8(const Text& t, const DevPattern& P, const Tables& T, uint64_t pos,
                                                     uint64_t first8)
{
  uint32_t state = 0;
#pragma unroll 1
  for (int i = 0; i < 8; ++i)
  {
    const uint32_t ch = static_cast<uint32_t>(first8) & 0xffu;
    first8 >>= 8;
    const uint32_t nxt = T.next[state * P.ncls + T.cls[ch]];
    if (nxt >= P.first_acc)
    {
      if (nxt == D_DEAD)
        return false;
      if (nxt < P.first_leaf)
        return true;
      return (__ldg(P.accept + nxt) & 0x7fffffffu) != 0;
    }
    if (nxt == 0 && P.acc0)
      return true;
    state = nxt;
  }
  // longer than 8 bytes: continue from text memory
  uint64_t p = pos + 8;
  for (;;)
  {
    if (p >= t.end)
      return false;
    const uint32_t ch = t.raw(p++);
    const uint32_t nxt = T.next[state * P.ncls + T.cls[ch]];
    if (nxt >= P.first_acc)
    {
      if (nxt == D_DEAD)
        return false;
      if (nxt < P.first_leaf)
        return true;
      return (__ldg(P.accept + nxt) & 0x7fffffffu) != 0;
    }
    if (nxt == 0 && P.acc0)
      return true;
    state = nxt;
  }
}

template <int KIND>
__device__ __forceinline__ bool attempt_at(const Text& t, const DevPattern& P, const Tables& T, uint64_t pos)
{
  if (P.one)
    return true; // the candidate test was the exact literal (lib/matcher.cpp:71-83)
  if (KIND != SK_META)
    return attempt_table(t, P, T, pos);
  return attempt_meta(t, P, pos);
}

// stage 2 for one survivor: the exact candidate predicate, then one anchored attempt.  Not inlined: the span loop is
// instantiated 16 times (4 spans x watch/cruise x full/guarded) and must stay within the instruction cache.
// the needle + hashed-predictor routines on an interior position (pos + 12 <= end): the 8 bytes the predicate reads
// come from three aligned 32-bit loads instead of eight guarded byte loads
__device__ __forceinline__ bool cand_pin_pmh_interior(const Text& t, const DevPattern& P, const Tables& T, uint64_t pos,
                                                      uint64_t& first8)
{
  const uint8_t* p = t.b + pos;
  const uint32_t sh = (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p)) & 3u) * 8;
  const uint32_t* a = reinterpret_cast<const uint32_t*>(p - (sh >> 3));
  const uint32_t lo = __ldg(a), mid = __ldg(a + 1), hi = __ldg(a + 2);
  const uint32_t x = __funnelshift_r(lo, mid, sh), y = __funnelshift_r(mid, hi, sh);
  const uint64_t xy = (static_cast<uint64_t>(y) << 32) | x;
  first8 = xy;
  const uint32_t ca = static_cast<uint32_t>(xy >> (8 * P.lcp)) & 0xffu, cb = static_cast<uint32_t>(xy >> (8 * P.lcs)) & 0xffu;
  if (P.adv == UGX_ADV_PIN1_PMH)
  {
    if (ca != P.chr[0] || cb != P.chr[1])
      return false;
  }
  else if (!bit256(P.pin_a, ca) || !bit256(P.pin_b, cb))
    return false;
  return pmh_xy(T.pred, x, y, P.min);
}

// advance_pattern_pma on an interior position (pos + 7 <= end): predict_match PM4 on the 4 bytes at pos
__device__ __forceinline__ bool cand_pma_interior(const Text& t, const Tables& T, uint64_t pos)
{
  const uint8_t* p = t.b + pos;
  const uint32_t sh = (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p)) & 3u) * 8;
  const uint32_t* a = reinterpret_cast<const uint32_t*>(p - (sh >> 3));
  return pm4_x(T.pred, __funnelshift_r(__ldg(a), __ldg(a + 1), sh));
}

template <int KIND>
__device__ __noinline__ bool stage2(Text t, const DevPattern& P, Tables T, uint64_t pos, bool exact)
{
  if (!exact)
  {
    const bool pin_pmh = P.adv == UGX_ADV_PIN_PMH || P.adv == UGX_ADV_PIN1_PMH;
    if (pos + 12 <= t.end && (pin_pmh || P.adv == UGX_ADV_PMA))
    {
      if (pin_pmh)
      {
        uint64_t first8;
        if (!cand_pin_pmh_interior(t, P, T, pos, first8))
          return false;
        if (KIND != SK_META && !P.one)
          return attempt_table_first8(t, P, T, pos, first8);
      }
      else if (!cand_pma_interior(t, T, pos))
        return false;
    }
    else if (!cand(t, P, T, pos))
      return false;
  }
  return attempt_at<KIND>(t, P, T, pos);
}

// the same attempt for a start state that does not accept, transitions through a Stepper (device_pattern.cuh): LDS by
// 32-bit shared addresses when the table is staged, the pattern's constants in registers, the 8 register bytes unrolled
// and predicated — 9 instructions per transition (the generic form above: 21, with LD.E and the constants re-read from
// the parameter bank)
template <bool STAGED>
__device__ __forceinline__ bool attempt_step_first8(const Text& t, const DevPattern& P, const Stepper& step, uint32_t first_acc,
                                                    uint64_t pos, uint32_t lo, uint32_t hi)
{
  uint32_t state = 0, nxt = 0;
  bool done = false;
#pragma unroll
  for (int i = 0; i < 8; ++i)
  {
    if (!done)
    {
      nxt = step.template at<STAGED>(state, __byte_perm(i < 4 ? lo : hi, 0, 0x4440 + (i & 3)));
      done = nxt >= first_acc;
      state = nxt;
    }
  }
  if (!done)
  {
    // longer than 8 bytes: continue from text memory
    uint64_t p = pos + 8;
    for (;;)
    {
      if (p >= t.end)
        return false;
      nxt = step.template at<STAGED>(state, t.raw(p++));
      if (nxt >= first_acc)
        break;
      state = nxt;
    }
  }
  // dead, accepting, or a leaf (accepting unless it is a dead end): the first such state reached decides
  if (nxt == D_DEAD)
    return false;
  if (nxt < P.first_leaf)
    return true;
  return (__ldg(P.accept + nxt) & 0x7fffffffu) != 0;
}

// Stage 2 for a batch of survivors of one span, in rounds of 32 (one survivor per lane).  The first 8 text bytes of an
// attempt come from three aligned 32-bit loads (they also feed the candidate predicate); `skip_cand`: the survivors need
// no candidate test (they are exact candidates already, or DevPattern::covers proves that a position starting a match
// always passes it).  Inlined: the DFA kernels instantiate the span evaluation twice only (full / guarded regions, one
// span per block), and a call here costs 120 bytes of spills around it (measured: 326 -> 368 GB/s on config 2).
// (A single flattened loop — a lane that finishes takes its next survivor at once, as in span_scan.cu — was measured
// SLOWER here, 195 vs 306 GB/s on config 2: a refill is three global loads and the predicate, and every refill of a few
// lanes stalls the other lanes of the warp.)
constexpr uint32_t SC_QCAP = 128; // survivor queue entries per warp (64 measures the same, 256 is 3 % slower)

#ifndef UGX_DRAIN_ATTR
#define UGX_DRAIN_ATTR __forceinline__
#endif
template <int KIND>
__device__ UGX_DRAIN_ATTR void drain_queue(Text t, const DevPattern& P, Tables T, uint64_t sbase, const uint16_t* queue,
                                         uint32_t qn, uint32_t* succ, uint32_t lane, bool skip_cand, bool interior,
                                         bool staged)
{
  if (KIND == SK_META || P.one || !interior)
  {
    for (uint32_t base = 0; base < qn; base += 32)
    {
      if (base + lane < qn)
      {
        const uint32_t off = queue[base + lane];
        if (stage2<KIND>(t, P, T, sbase + off, skip_cand))
          atomicOr(&succ[off >> 5], 1u << (off & 31));
      }
      __syncwarp();
    }
    return;
  }
  const uint8_t* __restrict__ sp = t.b + sbase; // 16-byte aligned (a span start)
  const bool pin_pmh = P.adv == UGX_ADV_PIN_PMH || P.adv == UGX_ADV_PIN1_PMH, pma = P.adv == UGX_ADV_PMA;
  const Stepper step(T, P.ncls, staged);
  const uint32_t first_acc = P.first_acc;
  const bool acc0 = P.acc0 != 0;
  for (uint32_t base = 0; base < qn; base += 32)
  {
    if (base + lane < qn)
    {
      const uint32_t off = queue[base + lane];
      const uint32_t sh = (off & 3u) * 8;
      const uint32_t* a = reinterpret_cast<const uint32_t*>(sp + (off & ~3u));
      const uint32_t l0 = __ldg(a), l1 = __ldg(a + 1), l2 = __ldg(a + 2);
      const uint32_t lo = __funnelshift_r(l0, l1, sh), hi = __funnelshift_r(l1, l2, sh);
      bool ok = true;
      if (!skip_cand)
      {
        if (pin_pmh)
        {
          const uint64_t xy = (static_cast<uint64_t>(hi) << 32) | lo;
          const uint32_t ca = static_cast<uint32_t>(xy >> (8 * P.lcp)) & 0xffu, cb = static_cast<uint32_t>(xy >> (8 * P.lcs)) & 0xffu;
          if (P.adv == UGX_ADV_PIN1_PMH)
            ok = ca == P.chr[0] && cb == P.chr[1];
          else
            ok = bit256(P.pin_a, ca) && bit256(P.pin_b, cb);
          ok = ok && pmh_xy(T.pred, lo, hi, P.min);
        }
        else if (pma)
          ok = pm4_x(T.pred, lo);
        else
          ok = cand(t, P, T, sbase + off);
      }
      if (ok)
      {
        const bool hit = acc0      ? attempt_table_first8(t, P, T, sbase + off, (static_cast<uint64_t>(hi) << 32) | lo)
                         : staged ? attempt_step_first8<true>(t, P, step, first_acc, sbase + off, lo, hi)
                                  : attempt_step_first8<false>(t, P, step, first_acc, sbase + off, lo, hi);
        if (hit)
          atomicOr(&succ[off >> 5], 1u << (off & 31));
      }
    }
    __syncwarp(); // (also: lanes must have reconverged before this function returns — the caller's own __syncwarp() after
                  // the call was measured NOT to order a late lane's atomic before an early lane's read of `succ`)
  }
}

// exact per-position candidates of one chunk (families without a byte-set plan, spans near the end of the buffer)
__device__ __noinline__ uint32_t exact_chunk_candidates(Text t, const DevPattern& P, Tables T, uint64_t base, uint32_t w0,
                                                        uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5)
{
  if (base >= t.end)
    return 0;
  if (base + 24 > t.end)
    return chunk_cand_generic(t, P, T, base);
  Window W;
  W.w[0] = w0;
  W.w[1] = w1;
  W.w[2] = w2;
  W.w[3] = w3;
  W.w[4] = w4;
  W.w[5] = w5;
  W.w[6] = 0;
  return chunk_cand_fast(W, t, P, T, base);
}

// Span evaluation for DFA patterns, in three steps:
//   1. survivors: the FilterPlan's byte-set terms (one shared-memory lookup per text byte, eight positions
//      accumulated per register; the right halo's fail bits come from the next lane) — a superset of the
//      reference's candidates.  Families without a byte-set plan, and spans near the end of the buffer, take the
//      exact per-position predicate instead;
//   2. the warp's survivors are compacted into a small shared-memory queue and handed out one per lane, so that
//      all 32 lanes run stage 2 whatever the survivors' distribution over chunks;
//   3. per survivor: the exact candidate predicate cand() (device_pattern.cuh) and one anchored DFA attempt.
template <int KIND>
struct DfaEval {
  const DevPattern& P;
  Tables T;
  Text t;
  uint32_t lane, next_lane;
  const uint32_t* lut;  // [256] shared: bit 8*t set = the byte fails term t
  const uint32_t* h4;   // [4096] shared: bit 8*t set = pmh[g] fails hashed-predictor step 3 + t; nullptr = unused
  const uint32_t* h4x;  // [4096] shared: the same storage when it holds the PM4 pair planes
  uint32_t h4_shift;
  bool pm2;             // lut/h4 hold the PM4 two-byte planes instead (q7 q6 by byte, q5 q4 by pair hash)
  uint16_t* queue;      // [SC_QCAP] per warp
  uint32_t* succ;       // [16] per warp: success bits of the span, bit (16 * lane + k)
  uint32_t nterms, off0, off1, off2;
  bool use_lut;
  ViaTables via;        // k-gram viability of an attempt: a survivor that cannot start a match is dropped before stage 2
  mutable uint32_t via_tick = 0; // spans with survivors seen so far; every 64th measures what the table spares ...
  mutable bool via_keep = true;  // ... and the next 63 use it only if that is worth its two lookups per byte
  mutable bool via_first = false; // the table alone picks the survivors (it is at least as selective as the prefilter's
                                  // first-stage planes, which are then not evaluated at all; stage 2 stays exact)
  bool cover = false;             // DevPattern::covers: an interior survivor needs no candidate test
  bool staged = false;            // T.next is the shared-memory copy of the table

  __device__ __forceinline__ bool operator()(const uint32_t (&w)[7], uint64_t sbase, uint32_t& succ16) const
  {
    const bool interior = sbase + SC_SPAN + 24 <= t.end; // uniform
    uint32_t surv = 0;
    bool exact;
    if (via.on && interior && via_first && (via_tick & 63u) != 0)
    {
      ++via_tick;
      exact = false;
      Window W;
#pragma unroll
      for (int i = 0; i < 7; ++i)
        W.w[i] = w[i];
      surv = viable16(via, W);
      if (!__any_sync(0xffffffffu, surv != 0))
        return false;
    }
    else
    {
    if (h4 != nullptr && !pm2 && interior)
    {
      // hashed-predictor terms: one rolling 12-bit hash and one lookup per text byte.  Byte i of the predictor
      // window of position k is text byte k + h4_shift + i; c[] is the chunk shifted by h4_shift.
      exact = false;
      uint32_t c[6];
#pragma unroll
      for (int i = 0; i < 6; ++i)
        c[i] = __funnelshift_r(w[i], w[i + 1], 8 * h4_shift);
      uint32_t g = 0, acc_a = 0, acc_b = 0, acc_c = 0;
#pragma unroll
      for (int p = 0; p <= 20; ++p)
      {
        const uint32_t byte = (c[p >> 2] >> (8 * (p & 3))) & 0xffu;
        g = ((g << 3) ^ byte) & (UGX_HASH - 1);
        if (p >= 3)
        {
          const uint32_t v = h4[g];
          if (p <= 10)
            acc_a = acc_a * 2 + v;
          else if (p <= 18)
            acc_b = acc_b * 2 + v;
          else
            acc_c = acc_c * 2 + v;
        }
      }
      // acc_a: p = 3..10 (first at bit 7 of each plane byte), acc_b: p = 11..18, acc_c: p = 19, 20 (bits 1, 0).
      // Bit-reverse: plane t moves to byte 3 - t with p ascending from its bit 0.
      const uint32_t ra = __brev(acc_a), rb = __brev(acc_b), rc = __brev(acc_c << 6);
      // step 3 + t at position k reads p = k + 3 + t
      uint32_t fail = (ra >> 24) | ((rb >> 24) << 8);                                              // t = 0: k = p - 3
      fail |= (((ra >> 16) & 0xffu) >> 1) | (((rb >> 16) & 0xffu) << 7) | (((rc >> 16) & 1u) << 15); // t = 1: k = p - 4
      fail |= (((ra >> 8) & 0xffu) >> 2) | (((rb >> 8) & 0xffu) << 6) | (((rc >> 8) & 3u) << 14);    // t = 2: k = p - 5
      surv = ~fail & 0xffffu;
    }
    else if (pm2 && interior)
    {
      // PM4 two-byte term: per position one lookup by byte (q7, q6) and one by the pair hash (q5, q4)
      exact = false;
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int k = 7; k >= 0; --k)
      {
        const uint32_t b0 = (w[k >> 2] >> (8 * (k & 3))) & 0xffu, b1 = (w[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xffu;
        const uint32_t d0 = (w[(k + 8) >> 2] >> (8 * (k & 3))) & 0xffu, d1 = (w[(k + 9) >> 2] >> (8 * ((k + 1) & 3))) & 0xffu;
        lo = lo * 2 + (lut[b0] | h4x[(b0 << 3) ^ b1]);
        hi = hi * 2 + (lut[d0] | h4x[(d0 << 3) ^ d1]);
      }
      const uint32_t q7 = __byte_perm(lo, hi, 0x4440), q6 = __byte_perm(lo, hi, 0x4451);
      const uint32_t q5 = __byte_perm(lo, hi, 0x4462), q4 = __byte_perm(lo, hi, 0x4473);
      surv = ~(q7 & q5 & (q4 | q6)) & 0xffffu;
    }
    else if (use_lut && interior)
    {
      exact = false;
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int k = 7; k >= 0; --k)
      {
        lo = lo * 2 + lut[(w[k >> 2] >> (8 * (k & 3))) & 0xffu];
        hi = hi * 2 + lut[(w[2 + (k >> 2)] >> (8 * (k & 3))) & 0xffu];
      }
      // fail masks of this chunk: f01 = term 0 (bits 0-15) | term 1 (bits 16-31), f2 = term 2
      const uint32_t f01 = __byte_perm(lo, hi, 0x5140);
      const uint32_t f2 = __byte_perm(lo, hi, 0x4462) & 0xffffu;
      uint32_t n01 = __shfl_sync(0xffffffffu, f01, next_lane);
      uint32_t n2 = nterms > 2 ? __shfl_sync(0xffffffffu, f2, next_lane) : 0u;
      if (lane == 31)
      {
        // the next span's first chunk is not evaluated yet: let its positions pass (stage 2 is exact)
        n01 = 0;
        n2 = 0;
      }
      uint32_t fail = ((f01 & 0xffffu) | (n01 << 16)) >> off0;
      if (nterms > 1)
        fail |= ((f01 >> 16) | (n01 & 0xffff0000u)) >> off1;
      if (nterms > 2)
        fail |= (f2 | (n2 << 16)) >> off2;
      surv = ~fail & 0xffffu;
    }
    else
    {
      exact = true;
      surv = exact_chunk_candidates(t, P, T, sbase + lane * 16, w[0], w[1], w[2], w[3], w[4], w[5]);
    }
    if (!__any_sync(0xffffffffu, surv != 0))
      return false;
    if (via.on && interior)
    {
      const bool probe = (via_tick++ & 63u) == 0;
      if (probe || via_keep)
      {
        Window W;
#pragma unroll
        for (int i = 0; i < 7; ++i)
          W.w[i] = w[i];
        const uint32_t v = viable16(via, W);
        if (probe)
        {
          const uint32_t spared = __reduce_add_sync(0xffffffffu, __popc(surv & ~v));
          const uint32_t extra = __reduce_add_sync(0xffffffffu, __popc(v & ~surv));
          via_keep = spared >= 32u;
          via_first = via_keep && !exact && extra <= 16u; // the table alone would hand stage 2 at most 16 more
        }
        surv &= v;
        if (!__any_sync(0xffffffffu, surv != 0))
          return false;
      }
    }
    }
    // ---- compaction + stage 2.  Per pass: a warp scan of the lanes' survivor counts; the lanes whose survivors fit
    // into the free part of the queue write ALL of them; the queue is drained (drain_queue) when a lane is left over and
    // at the end — one pass and one drain for all but the densest spans.
    const bool skip_cand = exact || (cover && interior);
    uint32_t qn = 0;
    bool more;
    do
    {
      const uint32_t cnt = __popc(surv);
      uint32_t incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<uint32_t>(d))
          incl += y;
      }
      const bool fits = incl <= SC_QCAP - qn; // monotone over lanes: a prefix of the lanes fits
      if (fits)
      {
        uint32_t at = qn + incl - cnt;
        while (surv != 0)
        {
          const uint32_t k = __ffs(surv) - 1;
          surv &= surv - 1;
          queue[at++] = static_cast<uint16_t>(lane * 16 + k);
        }
      }
      const uint32_t fitmask = __ballot_sync(0xffffffffu, fits);
      // entries written = inclusive count of the last fitting lane (an empty queue takes at least lane 0: cnt <= 16)
      if (fitmask != 0)
        qn += __shfl_sync(0xffffffffu, incl, 31u - __clz(fitmask));
      more = fitmask != 0xffffffffu;
      __syncwarp();
      drain_queue<KIND>(t, P, T, sbase, queue, qn, succ, lane, skip_cand, interior, staged);
      qn = 0;
      __syncwarp();
    } while (more);
    const uint32_t word = succ[lane >> 1];
    __syncwarp();
    if (lane < 16)
      succ[lane] = 0;
    __syncwarp();
    succ16 = (word >> (16 * (lane & 1))) & 0xffffu;
    return __any_sync(0xffffffffu, succ16 != 0);
  }
};

} // namespace

template <int KIND, bool WANT_NL, int THREADS>
#ifndef UGX_DFA_MINB
#define UGX_DFA_MINB 3
#endif
__global__ void __launch_bounds__(THREADS, THREADS >= 512 ? 1 : UGX_DFA_MINB)
count_lines_stream_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n, StreamArgs a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr uint32_t NWARPS = THREADS / 32;
  uint8_t* s_cls = smem;
  uint8_t* s_pred = s_cls + 256;
  uint8_t* s_tap = s_pred + UGX_HASH;
  uint32_t* s_lut = reinterpret_cast<uint32_t*>(s_tap + UGX_BTAP);
  uint32_t* s_succ = s_lut + 256;
  uint16_t* s_queue = reinterpret_cast<uint16_t*>(s_succ + 16 * NWARPS);
  uint32_t* s_h4 = reinterpret_cast<uint32_t*>(s_queue + SC_QCAP * NWARPS);
  uint8_t* s_via = reinterpret_cast<uint8_t*>(s_h4 + (a.use_h4 ? UGX_HASH : 0));
  uint16_t* s_next = reinterpret_cast<uint16_t*>(s_via + (a.use_via ? via_smem_bytes(P) : 0));
  // ---- tables -> shared memory by bulk asynchronous copies (cp.async.bulk, the TMA path without a tensor map)
  __shared__ __align__(8) uint64_t s_bar;
  stage_tables_bulk(&s_bar, s_cls, P.cls, s_pred, P.pred, s_tap, P.tap, s_next, P.next,
                    a.stage_table ? ((P.table_bytes + 15) / 16) * 16 : 0);
  const ViaTables via = via_stage(P, s_via, a.use_via != 0 && KIND == SK_TABLE);
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x)
    s_lut[i] = P.plan.pm2 ? (((__ldg(P.pred + i) >> 7) & 1u) | (((__ldg(P.pred + i) >> 6) & 1u) << 8)) : P.plan.lut[i];
  for (uint32_t i = threadIdx.x; i < 16 * NWARPS; i += blockDim.x)
    s_succ[i] = 0;
  if (a.use_h4)
    for (uint32_t i = threadIdx.x; i < UGX_HASH; i += blockDim.x)
    {
      // missing terms (h4_terms < 3) never fail
      const uint32_t e = __ldg(P.pred + i);
      uint32_t v = 0;
      if (P.plan.pm2)
        v = (((e >> 5) & 1u) << 16) | (((e >> 4) & 1u) << 24);
      else
        for (uint32_t tt = 0; tt < P.plan.h4_terms; ++tt)
          v |= ((e >> (3 + tt)) & 1u) << (8 * tt);
      s_h4[i] = v;
    }
  __syncthreads();
  Tables T;
  T.cls = s_cls;
  T.pred = s_pred;
  T.tap = s_tap;
  T.next = a.stage_table ? s_next : P.next;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  DfaEval<KIND> ev{P, T, Text{buf, n}, lane, (lane + 1) & 31, s_lut, a.use_h4 ? s_h4 : nullptr, s_h4, P.plan.h4_shift,
                   P.plan.pm2 != 0 && a.use_h4 != 0,
                   s_queue + SC_QCAP * wid, s_succ + 16 * wid,
                   P.plan.nterms, P.plan.t_off[0], P.plan.t_off[1], P.plan.t_off[2],
                   P.plan.kind == FK_LUT && P.plan.nterms >= 1, via};
  ev.cover = P.covers != 0 && a.no_cover == 0 && KIND == SK_TABLE;
  ev.staged = a.stage_table != 0;
  stream_scan<WANT_NL, false, 1>(buf, n, a, ev);
}

static bool stream_use_h4(const DevPattern& P) { return P.plan.h4_terms >= 1 || P.plan.pm2 != 0; }

static bool stream_use_via(const DevPattern& P) { return P.via_k != 0 && P.has_meta == 0 && P.one == 0 && via_smem_bytes(P) <= 40 * 1024; }

static size_t stream_smem_bytes(const DevPattern& P, bool stage, int threads)
{
  return 256 + UGX_HASH + UGX_BTAP + 1024 + (threads / 32) * (64 + 2 * SC_QCAP) + (stream_use_h4(P) ? 4 * UGX_HASH : 0) +
         (stream_use_via(P) ? via_smem_bytes(P) : 0) + (stage ? ((P.table_bytes + 15) / 16) * 16 : 0);
}

bool count_lines_stream_eligible(const DevPattern& P)
{
  return P.lbk == 0 && (P.flags & UGX_OPT_W) == 0 && P.adv != UGX_ADV_NONE;
}

uint64_t stream_regions(uint64_t n) { return (n + SC_REGION - 1) / SC_REGION; }

int stream_grid(uint64_t n, int sm_count, int per_sm, int threads)
{
  if (per_sm < 1)
    per_sm = 1;
  uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;
  const uint64_t warps = threads / 32;
  const uint64_t need = (stream_regions(n) + warps - 1) / warps;
  if (g > need)
    g = need;
  if (g == 0)
    g = 1;
  if (g > STREAM_MAX_GRID)
    g = STREAM_MAX_GRID;
  return static_cast<int>(g);
}

#ifndef UGX_BIG_THREADS
#define UGX_BIG_THREADS 1024 // a big staged table leaves room for one CTA per SM
#endif

template <int KIND, bool WANT_NL, int THREADS>
static cudaError_t launch_dfa(const DevPattern& P, const uint8_t* buf, uint64_t n, const StreamArgs& a, size_t smem,
                              int sm_count, cudaStream_t st)
{
  auto kern = count_lines_stream_kernel<KIND, WANT_NL, THREADS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UGX_MAX_DYN_SMEM);
  if (e != cudaSuccess)
    return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
  if (e != cudaSuccess)
    return e;
  kern<<<stream_grid((a.region_end - a.region_begin) * SC_REGION, sm_count, per_sm, THREADS), THREADS, smem, st>>>(P, buf, n, a);
  return cudaGetLastError();
}

cudaError_t launch_count_lines_stream(const DevPattern& P, const uint8_t* buf, uint64_t n, StreamArgs a, bool want_nl,
                                      int sm_count, cudaStream_t st)
{
  if (count_lines_literal_eligible(P))
    return launch_count_lines_literal(P, buf, n, a, want_nl, sm_count, st);
  const bool meta = P.has_meta != 0;
  const bool stage = !meta && stream_smem_bytes(P, true, 1024) <= static_cast<size_t>(UGX_MAX_DYN_SMEM);
  // a big staged table leaves room for one CTA per SM: make it a full 1024-thread CTA
  const bool big = stage && stream_smem_bytes(P, true, 256) > 100 * 1024;
  const size_t smem = stream_smem_bytes(P, stage, big ? UGX_BIG_THREADS : 256);
  a.stage_table = stage ? 1u : 0u;
  a.use_h4 = stream_use_h4(P) ? 1u : 0u;
  a.use_via = stream_use_via(P) ? 1u : 0u;
  if (meta)
    return want_nl ? launch_dfa<SK_META, true, 256>(P, buf, n, a, smem, sm_count, st)
                   : launch_dfa<SK_META, false, 256>(P, buf, n, a, smem, sm_count, st);
  if (big)
    return want_nl ? launch_dfa<SK_TABLE, true, UGX_BIG_THREADS>(P, buf, n, a, smem, sm_count, st)
                   : launch_dfa<SK_TABLE, false, UGX_BIG_THREADS>(P, buf, n, a, smem, sm_count, st);
  return want_nl ? launch_dfa<SK_TABLE, true, 256>(P, buf, n, a, smem, sm_count, st)
                 : launch_dfa<SK_TABLE, false, 256>(P, buf, n, a, smem, sm_count, st);
}

} // namespace ugx
