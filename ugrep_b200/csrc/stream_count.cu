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
#include "tile_phase_a.cuh"

namespace ugx {

namespace {

constexpr int SC_SPANS = 4;                              // spans per block
constexpr uint32_t SC_SPAN = 512;                        // bytes per span: 32 lanes x 16
constexpr uint32_t SC_BLOCK = SC_SPANS * SC_SPAN;        // 2 KiB per warp iteration

enum StreamKind { SK_LITERAL = 0, SK_TABLE = 1, SK_META = 2 };

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

struct Blk {
  uint4 v[SC_SPANS];
  uint2 halo; // the 8 bytes after the block
};

__device__ __forceinline__ uint4 ldg128(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ uint32_t load_bytes(const uint8_t* __restrict__ buf, uint64_t n, uint64_t at)
{
  uint32_t x = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (at + b < n)
      x |= static_cast<uint32_t>(__ldg(buf + at + b)) << (8 * b);
  return x;
}

__device__ __forceinline__ void load_block(Blk& B, const uint8_t* __restrict__ buf, uint64_t n, uint64_t bbase, uint32_t lane)
{
  if (bbase + SC_BLOCK + 8 <= n)
  {
#pragma unroll
    for (int j = 0; j < SC_SPANS; ++j)
      B.v[j] = ldg128(buf + bbase + j * SC_SPAN + lane * 16);
    B.halo = __ldg(reinterpret_cast<const uint2*>(buf + bbase + SC_BLOCK));
  }
  else
  {
    // the last block(s) of the buffer: bytes at or past n read as zero
#pragma unroll
    for (int j = 0; j < SC_SPANS; ++j)
    {
      const uint64_t base = bbase + j * SC_SPAN + lane * 16;
      if (base + 16 <= n)
        B.v[j] = ldg128(buf + base);
      else
        B.v[j] = make_uint4(load_bytes(buf, n, base), load_bytes(buf, n, base + 4), load_bytes(buf, n, base + 8),
                            load_bytes(buf, n, base + 12));
    }
    B.halo = make_uint2(load_bytes(buf, n, bbase + SC_BLOCK), load_bytes(buf, n, bbase + SC_BLOCK + 4));
  }
}

// ---- literal first stage: two bytes of the literal at fixed offsets (FilterPlan FK_ANCHOR2) ------------
// Anchor 0 is the literal's first byte (no shift); anchor 1 sits Q1 words + sh1 bits further on.
struct Anchors {
  uint32_t c0, c1, sh1;
};

// per word: zero byte <=> both anchor bytes match at that position
template <int Q1>
__device__ __forceinline__ uint32_t anchor_word(const uint32_t (&w)[7], int i, const Anchors& A)
{
  return (w[i] ^ A.c0) | (__funnelshift_r(w[i + Q1], w[i + Q1 + 1], A.sh1) ^ A.c1);
}

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

// ---- per-warp line state -----------------------------------------------------------------------------
struct LineState {
  uint32_t cin;      // the line open at the cursor already has a success (starts at 1: the region's head line is deferred)
  bool seen_nl;      // a newline was seen in this region
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
}

__device__ __forceinline__ uint32_t newline_mask_exact(const uint32_t (&w)[7])
{
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    m |= (((zero_bytes(w[i] ^ 0x0a0a0a0au)) * 0x00204081u) >> 28) << (4 * i);
  return m;
}

} // namespace

template <int KIND, bool WANT_NL, int Q1>
__global__ void __launch_bounds__(STREAM_THREADS, KIND == SK_LITERAL ? 3 : 2)
count_lines_stream_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n,
                          StreamArgs a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ unsigned long long s_red[2 * (STREAM_THREADS / 32)];
  __shared__ uint32_t s_last;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  Tables T;
  T.cls = nullptr;
  T.next = nullptr;
  T.pred = nullptr;
  T.tap = nullptr;
  if (KIND != SK_LITERAL)
  {
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
    T.cls = s_cls;
    T.pred = s_pred;
    T.tap = s_tap;
    T.next = a.stage_table ? s_next : P.next;
  }
  const Text t{buf, n};
  Anchors A;
  A.c0 = P.plan.a_chr[0];
  A.c1 = P.plan.a_chr[1];
  A.sh1 = (P.plan.a_off[1] & 3) * 8;

  const uint64_t nregions = (n + SC_REGION - 1) / SC_REGION;
  unsigned long long my_lines = 0, my_newlines = 0;
  uint32_t warp_uniform_lines = 0;

  // ---- regions from the ticket counter; the first block of the next region is requested before the
  //      current region's last block is evaluated
  uint64_t r = 0;
  {
    unsigned long long tk = 0;
    if (lane == 0)
      tk = atomicAdd(a.ticket, 1ull);
    r = __shfl_sync(0xffffffffu, tk, 0);
  }
  Blk cur;
  if (r < nregions)
    load_block(cur, buf, n, r * SC_REGION, lane);
  while (r < nregions)
  {
    const uint64_t rbase = r * SC_REGION;
    const uint64_t rend = rbase + SC_REGION < n ? rbase + SC_REGION : n;
    const uint32_t nblocks = static_cast<uint32_t>((rend - rbase + SC_BLOCK - 1) / SC_BLOCK);
    uint64_t rnext = 0;
    {
      unsigned long long tk = 0;
      if (lane == 0)
        tk = atomicAdd(a.ticket, 1ull);
      rnext = __shfl_sync(0xffffffffu, tk, 0);
    }
    LineState L;
    L.cin = 1;
    L.seen_nl = false;
    L.head = false;
    L.ucount = 0;
    L.lcount = 0;
    uint32_t nlacc = 0; // WANT_NL: per-byte-lane newline counters (at most 128 per region)
    for (uint32_t b = 0; b < nblocks; ++b)
    {
      const uint64_t bbase = rbase + static_cast<uint64_t>(b) * SC_BLOCK;
      Blk nxt;
      const bool more = b + 1 < nblocks || rnext < nregions;
      if (more)
        load_block(nxt, buf, n, b + 1 < nblocks ? bbase + SC_BLOCK : rnext * SC_REGION, lane);
#pragma unroll
      for (int j = 0; j < SC_SPANS; ++j)
      {
        const uint64_t sbase = bbase + j * SC_SPAN;
        if (sbase >= n)
          break;
        const uint64_t base = sbase + lane * 16;
        uint32_t w[7];
        w[0] = cur.v[j].x;
        w[1] = cur.v[j].y;
        w[2] = cur.v[j].z;
        w[3] = cur.v[j].w;
        // halo: the first 8 bytes of the next chunk live in the next lane; lane 31 takes lane 0's next span
        const uint32_t nx = j + 1 < SC_SPANS ? cur.v[j + 1 < SC_SPANS ? j + 1 : j].x : cur.halo.x;
        const uint32_t ny = j + 1 < SC_SPANS ? cur.v[j + 1 < SC_SPANS ? j + 1 : j].y : cur.halo.y;
        w[4] = __shfl_sync(0xffffffffu, lane == 0 ? nx : w[0], (lane + 1) & 31);
        w[5] = __shfl_sync(0xffffffffu, lane == 0 ? ny : w[1], (lane + 1) & 31);
        w[6] = 0;
        // ---- newline test
        uint32_t nl_any;
        uint32_t nl_exact[4];
        if (WANT_NL)
        {
          nl_any = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
          {
            nl_exact[i] = zero_bytes(w[i] ^ 0x0a0a0a0au);
            nlacc += nl_exact[i] >> 7;
            nl_any |= nl_exact[i];
          }
        }
        else
        {
          nl_any = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            nl_any = zero_any(w[i] ^ 0x0a0a0a0au, nl_any);
          nl_any &= 0x80808080u;
        }
        // ---- successes
        uint32_t succ16 = 0;
        if (KIND == SK_LITERAL)
        {
          uint32_t acc = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc = zero_any(anchor_word<Q1>(w, i, A), acc);
          if (__any_sync(0xffffffffu, (acc & 0x80808080u) != 0))
          {
            if ((acc & 0x80808080u) != 0)
            {
              uint32_t surv = 0;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                surv |= ((zero_bytes(anchor_word<Q1>(w, i, A)) * 0x00204081u) >> 28) << (4 * i);
              while (surv != 0)
              {
                const uint32_t k = __ffs(surv) - 1;
                surv &= surv - 1;
                if (literal_at(t, P, base + k))
                  succ16 |= 1u << k;
              }
            }
          }
        }
        else
        {
          uint32_t cm = 0;
          if (base < n)
          {
            if (base + 24 <= n)
            {
              Window W;
#pragma unroll
              for (int i = 0; i < 7; ++i)
                W.w[i] = w[i];
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
        }
        // ---- lines
        if (__any_sync(0xffffffffu, succ16 != 0))
        {
          resolve_span(L, newline_mask_exact(w), succ16);
        }
        else if (__any_sync(0xffffffffu, nl_any != 0))
        {
          L.seen_nl = true;
          L.cin = 0;
        }
      }
      if (more)
        cur = nxt;
    }
    // ---- publish the region: bit 0 has newline, bit 1 head success, bit 2 carry out
    if (lane == 0)
    {
      const uint32_t g = L.seen_nl ? L.cin : (L.head ? 1u : 0u);
      a.region_sum[a.first_region + r] = static_cast<uint8_t>((L.seen_nl ? 1u : 0u) | (L.head ? 2u : 0u) | (g << 2));
    }
    my_lines += L.lcount;
    warp_uniform_lines += L.ucount;
    if (WANT_NL)
    {
      const uint32_t pair = (nlacc & 0x00ff00ffu) + ((nlacc >> 8) & 0x00ff00ffu);
      my_newlines += (pair & 0xffffu) + (pair >> 16);
    }
    r = rnext;
  }
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
  __shared__ uint8_t s_slice[STREAM_THREADS];
  if (a.finalize)
  {
    // thread i chains the slice [lo, hi) assuming no carry in; the slices are then chained serially
    const uint64_t total = a.first_region + nregions;
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

static size_t stream_smem_bytes(const DevPattern& P, int kind, bool stage)
{
  if (kind == SK_LITERAL)
    return 0;
  return 256 + UGX_HASH + UGX_BTAP + (stage ? ((P.table_bytes + 15) / 16) * 16 : 0);
}

bool count_lines_stream_eligible(const DevPattern& P)
{
  return P.lbk == 0 && (P.flags & UGX_OPT_W) == 0 && P.adv != UGX_ADV_NONE;
}

uint64_t stream_regions(uint64_t n) { return (n + SC_REGION - 1) / SC_REGION; }

cudaError_t launch_count_lines_stream(const DevPattern& P, const uint8_t* buf, uint64_t n, StreamArgs a, bool want_nl,
                                      int sm_count, cudaStream_t st, int* grid_out)
{
  const int kind = (P.one && P.adv == UGX_ADV_STRING && P.plan.kind == FK_ANCHOR2 && P.plan.a_off[0] == 0 && P.plan.a_off[1] <= 8) ? SK_LITERAL : P.has_meta ? SK_META : SK_TABLE;
  const bool stage = kind == SK_TABLE && stream_smem_bytes(P, kind, true) <= 227 * 1024 - 2048;
  const size_t smem = stream_smem_bytes(P, kind, stage);
  a.stage_table = stage ? 1u : 0u;
  int per_sm = 1;
  cudaError_t e = cudaSuccess;
  const uint32_t q1 = P.plan.a_off[1] >> 2;
#define UGX_STREAM_LAUNCH(K, NLF, Q)                                                                                       \
  do                                                                                                                        \
  {                                                                                                                         \
    auto kern = count_lines_stream_kernel<K, NLF, Q>;                                                                       \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));                    \
    if (e != cudaSuccess)                                                                                                   \
      return e;                                                                                                             \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STREAM_THREADS, smem);                                 \
    if (e != cudaSuccess)                                                                                                   \
      return e;                                                                                                             \
    if (per_sm < 1)                                                                                                         \
      per_sm = 1;                                                                                                           \
    uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;                                                                  \
    const uint64_t need = (stream_regions(n) + STREAM_THREADS / 32 - 1) / (STREAM_THREADS / 32);                            \
    if (g > need)                                                                                                           \
      g = need;                                                                                                             \
    if (g == 0)                                                                                                             \
      g = 1;                                                                                                                \
    if (g > STREAM_MAX_GRID)                                                                                                \
      g = STREAM_MAX_GRID;                                                                                                  \
    if (grid_out)                                                                                                           \
      *grid_out = static_cast<int>(g);                                                                                      \
    kern<<<static_cast<int>(g), STREAM_THREADS, smem, st>>>(P, buf, n, a);                                                  \
  } while (0)
#define UGX_STREAM_LIT(NLF)                                                                                                \
  do                                                                                                                        \
  {                                                                                                                         \
    if (q1 == 0)                                                                                                            \
      UGX_STREAM_LAUNCH(SK_LITERAL, NLF, 0);                                                                                \
    else if (q1 == 1)                                                                                                       \
      UGX_STREAM_LAUNCH(SK_LITERAL, NLF, 1);                                                                                \
    else                                                                                                                    \
      UGX_STREAM_LAUNCH(SK_LITERAL, NLF, 2);                                                                                \
  } while (0)
  if (kind == SK_LITERAL)
  {
    if (want_nl)
      UGX_STREAM_LIT(true);
    else
      UGX_STREAM_LIT(false);
  }
  else if (kind == SK_META)
  {
    if (want_nl)
      UGX_STREAM_LAUNCH(SK_META, true, 0);
    else
      UGX_STREAM_LAUNCH(SK_META, false, 0);
  }
  else
  {
    if (want_nl)
      UGX_STREAM_LAUNCH(SK_TABLE, true, 0);
    else
      UGX_STREAM_LAUNCH(SK_TABLE, false, 0);
  }
#undef UGX_STREAM_LIT
#undef UGX_STREAM_LAUNCH
  return cudaGetLastError();
}

} // namespace ugx
