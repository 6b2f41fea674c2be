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
#include "stream_common.cuh"
#include "tile_phase_a.cuh"

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

template <int KIND>
__device__ __forceinline__ bool attempt_at(const Text& t, const DevPattern& P, const Tables& T, uint64_t pos)
{
  if (P.one)
    return true; // the candidate test was the exact literal (lib/matcher.cpp:71-83)
  if (KIND != SK_META)
    return attempt_table(t, P, T, pos);
  return attempt_meta(t, P, pos);
}

// candidates of the chunk by the reference's prefilter predicate, then one anchored attempt per candidate
template <int KIND>
struct DfaEval {
  const DevPattern& P;
  Tables T;
  Text t;
  uint32_t lane;

  __device__ __forceinline__ bool operator()(const uint32_t (&w)[7], uint64_t sbase, uint32_t& succ16) const
  {
    const uint64_t base = sbase + lane * 16;
    uint32_t cm = 0;
    if (base < t.end)
    {
      if (base + 24 <= t.end)
      {
        Window W;
#pragma unroll
        for (int i = 0; i < 6; ++i)
          W.w[i] = w[i];
        W.w[6] = 0;
        cm = chunk_cand_fast(W, t, P, T, base);
      }
      else
        cm = chunk_cand_generic(t, P, T, base);
    }
    while (cm != 0)
    {
      const uint32_t k = __ffs(cm) - 1;
      cm &= cm - 1;
      if (attempt_at<KIND>(t, P, T, base + k))
        succ16 |= 1u << k;
    }
    return __any_sync(0xffffffffu, succ16 != 0);
  }
};

} // namespace

template <int KIND, bool WANT_NL>
__global__ void __launch_bounds__(STREAM_THREADS, 2)
count_lines_stream_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n, StreamArgs a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* s_cls = smem;
  uint8_t* s_pred = s_cls + 256;
  uint8_t* s_tap = s_pred + UGX_HASH;
  uint16_t* s_next = reinterpret_cast<uint16_t*>(s_tap + UGX_BTAP);
  for (uint32_t i = threadIdx.x; i < 256 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(s_cls)[i] = __ldg(reinterpret_cast<const uint32_t*>(P.cls) + i);
  for (uint32_t i = threadIdx.x; i < UGX_HASH / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_pred)[i] = __ldg(reinterpret_cast<const uint4*>(P.pred) + i);
  for (uint32_t i = threadIdx.x; i < UGX_BTAP / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_tap)[i] = __ldg(reinterpret_cast<const uint4*>(P.tap) + i);
  if (a.stage_table)
    for (uint32_t i = threadIdx.x; i < (P.table_bytes + 15) / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(s_next)[i] = __ldg(reinterpret_cast<const uint4*>(P.next) + i);
  __syncthreads();
  Tables T;
  T.cls = s_cls;
  T.pred = s_pred;
  T.tap = s_tap;
  T.next = a.stage_table ? s_next : P.next;
  DfaEval<KIND> ev{P, T, Text{buf, n}, threadIdx.x & 31};
  stream_scan<WANT_NL>(buf, n, a, ev);
}

static size_t stream_smem_bytes(const DevPattern& P, bool stage)
{
  return 256 + UGX_HASH + UGX_BTAP + (stage ? ((P.table_bytes + 15) / 16) * 16 : 0);
}

bool count_lines_stream_eligible(const DevPattern& P)
{
  return P.lbk == 0 && (P.flags & UGX_OPT_W) == 0 && P.adv != UGX_ADV_NONE;
}

uint64_t stream_regions(uint64_t n) { return (n + SC_REGION - 1) / SC_REGION; }

int stream_grid(uint64_t n, int sm_count, int per_sm)
{
  if (per_sm < 1)
    per_sm = 1;
  uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;
  const uint64_t need = (stream_regions(n) + STREAM_THREADS / 32 - 1) / (STREAM_THREADS / 32);
  if (g > need)
    g = need;
  if (g == 0)
    g = 1;
  if (g > STREAM_MAX_GRID)
    g = STREAM_MAX_GRID;
  return static_cast<int>(g);
}

template <int KIND, bool WANT_NL>
static cudaError_t launch_dfa(const DevPattern& P, const uint8_t* buf, uint64_t n, const StreamArgs& a, size_t smem,
                              int sm_count, cudaStream_t st)
{
  auto kern = count_lines_stream_kernel<KIND, WANT_NL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess)
    return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STREAM_THREADS, smem);
  if (e != cudaSuccess)
    return e;
  kern<<<stream_grid(n, sm_count, per_sm), STREAM_THREADS, smem, st>>>(P, buf, n, a);
  return cudaGetLastError();
}

cudaError_t launch_count_lines_stream(const DevPattern& P, const uint8_t* buf, uint64_t n, StreamArgs a, bool want_nl,
                                      int sm_count, cudaStream_t st)
{
  if (count_lines_literal_eligible(P))
    return launch_count_lines_literal(P, buf, n, a, want_nl, sm_count, st);
  const bool meta = P.has_meta != 0;
  const bool stage = !meta && stream_smem_bytes(P, true) <= 227 * 1024 - 2048;
  const size_t smem = stream_smem_bytes(P, stage);
  a.stage_table = stage ? 1u : 0u;
  if (meta)
    return want_nl ? launch_dfa<SK_META, true>(P, buf, n, a, smem, sm_count, st)
                   : launch_dfa<SK_META, false>(P, buf, n, a, smem, sm_count, st);
  return want_nl ? launch_dfa<SK_TABLE, true>(P, buf, n, a, smem, sm_count, st)
                 : launch_dfa<SK_TABLE, false>(P, buf, n, a, smem, sm_count, st);
}

} // namespace ugx
