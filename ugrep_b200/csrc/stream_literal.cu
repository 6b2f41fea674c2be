// stream_literal.cu — `ugrep -c -F literal`: the streaming count (stream_common.cuh) with the literal
// prefilter of the reference's advance_string family (lib/matcher.cpp:3297-3549, the AVX2/AVX-512BW
// needle search of lib/matcher_avx2.cpp:78-186) restated for 32-bit SWAR lanes.
//
// A pattern that is one literal (Pattern::one_) matches exactly where the literal occurs
// (lib/matcher.cpp:71-83, 659-669), so "success at byte k" is "the literal starts at k".  First stage, every
// byte, branch-free: two bytes of the literal — its first byte and the rarest of bytes 1..12 (FilterPlan
// FK_ANCHOR2, pattern_host.cpp) — are compared at all 16 positions of a chunk at once:
//     x_i = (w_i ^ c0) | (w'_i ^ c1)          w' = the window shifted by the second anchor's offset
// has a zero byte exactly where both anchors match, and (x - 0x01010101) & ~x & 0x80808080 != 0 detects a zero
// byte; four words are folded before one vote.  Survivors (rare) are verified by the whole warp, one byte
// of the literal per lane.
#include "device_pattern.cuh"
#include "scan_kernels.hpp"
#include "stream_common.cuh"

#ifndef UGX_LIT_MINB
#define UGX_LIT_MINB 4
#endif

namespace ugx {

namespace {

template <int Q1, bool ALIGNED>
struct LiteralEval {
  const uint8_t* __restrict__ buf;
  uint64_t n;
  const uint8_t* chr; // the literal (kernel parameter space)
  uint32_t len;
  uint32_t c0, c1, sh1;
  uint32_t lane;

  __device__ __forceinline__ uint32_t anchor_word(const uint32_t (&w)[7], int i) const
  {
    const uint32_t second = ALIGNED ? w[i + Q1] : __funnelshift_r(w[i + Q1], w[i + Q1 + 1], sh1);
    return (w[i] ^ c0) | (second ^ c1);
  }

  __device__ __forceinline__ bool operator()(const uint32_t (&w)[7], uint64_t sbase, uint32_t& succ16) const
  {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      acc = zero_any(anchor_word(w, i), acc);
    const bool hit = (acc & 0x80808080u) != 0;
    uint32_t hitmask = __ballot_sync(0xffffffffu, hit);
    if (hitmask == 0)
      return false;
    // exact survivor positions of the lanes that hit
    uint32_t surv = 0;
    if (hit)
    {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        surv |= flags_to_nibble(zero_bytes(anchor_word(w, i))) << (4 * i);
    }
    // the warp verifies the survivors one at a time: lane l compares bytes l, l + 32, ... of the literal
    bool found = false;
    while (hitmask != 0)
    {
      const uint32_t src = __ffs(hitmask) - 1;
      hitmask &= hitmask - 1;
      uint32_t todo = __shfl_sync(0xffffffffu, surv, src);
      while (todo != 0)
      {
        const uint32_t k = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint64_t pos = sbase + src * 16 + k;
        if (pos + len > n)
          continue;
        bool differs = false;
        for (uint32_t i = lane; i < len; i += 32)
          differs |= __ldg(buf + pos + i) != chr[i];
        if (!__any_sync(0xffffffffu, differs))
        {
          found = true;
          if (lane == src)
            succ16 |= 1u << k;
        }
      }
    }
    return found;
  }
};

} // namespace

template <bool WANT_NL, int Q1, bool ALIGNED>
__global__ void __launch_bounds__(STREAM_THREADS, UGX_LIT_MINB)
count_lines_literal_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n, StreamArgs a)
{
  LiteralEval<Q1, ALIGNED> ev;
  ev.buf = buf;
  ev.n = n;
  ev.chr = P.chr;
  ev.len = P.len;
  ev.c0 = P.plan.a_chr[0];
  ev.c1 = P.plan.a_chr[1];
  ev.sh1 = (P.plan.a_off[1] & 3) * 8;
  ev.lane = threadIdx.x & 31;
  asm volatile("" : "+r"(ev.lane));
  stream_scan<WANT_NL, true, 4>(buf, n, a, ev);
}

bool count_lines_literal_eligible(const DevPattern& P)
{
  return P.one && P.adv == UGX_ADV_STRING && P.lbk == 0 && (P.flags & UGX_OPT_W) == 0 && P.plan.kind == FK_ANCHOR2 &&
         P.plan.a_off[0] == 0 && P.plan.a_off[1] <= 12 && P.len >= 2;
}

template <bool WANT_NL, int Q1, bool ALIGNED>
static cudaError_t launch_literal(const DevPattern& P, const uint8_t* buf, uint64_t n, const StreamArgs& a, int sm_count,
                                  cudaStream_t st)
{
  auto kern = count_lines_literal_kernel<WANT_NL, Q1, ALIGNED>;
  int per_sm = 1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STREAM_THREADS, 0);
  if (e != cudaSuccess)
    return e;
  kern<<<stream_grid((a.region_end - a.region_begin) * SC_REGION, sm_count, per_sm, STREAM_THREADS), STREAM_THREADS, 0, st>>>(P, buf, n, a);
  return cudaGetLastError();
}

cudaError_t launch_count_lines_literal(const DevPattern& P, const uint8_t* buf, uint64_t n, const StreamArgs& a, bool want_nl,
                                       int sm_count, cudaStream_t st)
{
  const uint32_t q1 = P.plan.a_off[1] >> 2;
  const bool aligned = (P.plan.a_off[1] & 3) == 0;
#define UGX_LIT(NLF)                                                                    \
  do                                                                                    \
  {                                                                                     \
    if (q1 == 0)                                                                        \
      return launch_literal<NLF, 0, false>(P, buf, n, a, sm_count, st);                 \
    if (q1 == 1)                                                                        \
      return aligned ? launch_literal<NLF, 1, true>(P, buf, n, a, sm_count, st)         \
                     : launch_literal<NLF, 1, false>(P, buf, n, a, sm_count, st);       \
    if (q1 == 2)                                                                        \
      return aligned ? launch_literal<NLF, 2, true>(P, buf, n, a, sm_count, st)         \
                     : launch_literal<NLF, 2, false>(P, buf, n, a, sm_count, st);       \
    return launch_literal<NLF, 3, true>(P, buf, n, a, sm_count, st);                    \
  } while (0)
  if (want_nl)
    UGX_LIT(true);
  UGX_LIT(false);
#undef UGX_LIT
}

} // namespace ugx
