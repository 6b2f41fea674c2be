// fast_kernels.cu — count_lines_any_kernel: `ugrep -c` for patterns without look-back.
//
// Without look-back (lbk_ == 0) and without option W, a line matches iff SOME candidate position in
// it starts a non-empty DFA match: the reference's find loop (lib/matcher.cpp:42-750) tries the
// candidates of a line in order until one succeeds, and `ugrep -c` then skips to the next line
// (src/ugrep.cpp:10567-10586).  That makes the whole job position-parallel:
//
//   phase A  prefilter every byte of the tile (tile_phase_a.cuh)        -> candidate + newline bitmaps
//   phase B  anchored DFA attempt at every candidate, balanced over the warp through a small
//            shared-memory queue                                           -> success bitmap
//   phase C  map each success to the start of its line (newline bitmap + a block-wide
//            "last newline so far" scan), OR into a line bitmap, popcount   -> matching lines
//
// A line belongs to the tile it starts in; the one line that runs past the end of the tile is
// followed by the whole CTA ("tail scan") until it matches or ends.
#include "device_pattern.cuh"
#include "ptx.cuh"
#include "line_match.cuh"
#include "scan_kernels.hpp"
#include "tile_phase_a.cuh"

namespace ugx {

namespace {

// does an anchored attempt at `pos` give a non-empty match?  (dense table: the first non-empty accept decides)
__device__ __forceinline__ bool attempt_table(const Text& t, const DevPattern& P, const Tables& T, uint64_t pos)
{
  uint32_t state = 0;
  uint64_t p = pos;
  for (;;)
  {
    const uint32_t acc = __ldg(P.accept + state);
    if ((acc & 0x7fffffffu) != 0 && p > pos)
      return true;
    if ((acc & 0x80000000u) || p >= t.end)
      return false;
    const uint32_t ch = t.raw(p++);
    const uint32_t nxt = T.next[state * P.ncls + T.cls[ch]];
    if (nxt == D_DEAD)
      return false;
    state = nxt;
  }
}

template <bool HAS_META>
__device__ __forceinline__ bool attempt(const Text& t, const DevPattern& P, const Tables& T, uint64_t pos)
{
  if (P.one)
    return true; // the candidate test was the exact literal (lib/matcher.cpp:71-83)
  if (!HAS_META)
    return attempt_table(t, P, T, pos);
  Cursor m;
  set_current(t, m, pos);
  m.txt = pos;
  m.len = 0;
  uint32_t retry = 0;
  run_dfa_opc(t, P, m, retry);
  return m.cap != 0 && m.cur > m.txt;
}

__device__ __forceinline__ int warp_incl_max(int v)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    int y = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= static_cast<uint32_t>(d))
      v = max(v, y);
  }
  return v;
}

} // namespace

template <bool HAS_META, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
count_lines_any_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n, uint64_t ntiles,
                       uint32_t stage_table, unsigned long long* __restrict__ totals)
{
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr uint32_t NW = ANY_TILE / 32; // bitmap words per tile
  __shared__ int s_warp_max[32];
  __shared__ int s_last_start;
  __shared__ uint32_t s_first_nl;
  __shared__ unsigned long long s_red[2];
  uint8_t* s_cls = smem;
  uint8_t* s_pred = s_cls + 256;
  uint8_t* s_tap = s_pred + UGX_HASH;
  uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_tap + UGX_BTAP);
  uint32_t* s_nl = s_cand + NW;
  uint32_t* s_succ = s_nl + NW;
  uint32_t* s_line = s_succ + NW;
  int* s_lastnl = reinterpret_cast<int*>(s_line + NW);
  uint16_t* s_queue = reinterpret_cast<uint16_t*>(s_lastnl + NW);
  uint16_t* s_next = s_queue + 32 * 64;
  // tables -> shared memory by bulk asynchronous copies (ptx.cuh)
  __shared__ __align__(8) uint64_t s_bar;
  stage_tables_bulk(&s_bar, s_cls, P.cls, s_pred, P.pred, s_tap, P.tap, s_next, P.next,
                    stage_table ? ((P.table_bytes + 15) / 16) * 16 : 0);
  if (threadIdx.x < 2)
    s_red[threadIdx.x] = 0;
  __syncthreads();
  Tables T;
  T.cls = s_cls;
  T.pred = s_pred;
  T.tap = s_tap;
  T.next = stage_table ? s_next : P.next;
  const Text t{buf, n};
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint16_t* my_queue = s_queue + wid * 64;
  uint32_t my_lines = 0, my_newlines = 0;

  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const uint64_t tile_base = tile * ANY_TILE;
    for (uint32_t wi = threadIdx.x; wi < NW; wi += blockDim.x)
    {
      s_succ[wi] = 0;
      s_line[wi] = 0;
    }
    // ---------------- phase A
    my_newlines += tile_phase_a<ANY_TILE / 16 / THREADS>(t, P, T, tile_base, reinterpret_cast<uint16_t*>(s_cand),
                                                         reinterpret_cast<uint16_t*>(s_nl));
    __syncthreads();
    // ---------------- phase B: every candidate gets one anchored attempt
    bool any_succ = false;
    if (P.one)
    {
      for (uint32_t wi = threadIdx.x; wi < NW; wi += blockDim.x)
      {
        const uint32_t c = s_cand[wi];
        s_succ[wi] = c;
        any_succ |= c != 0;
      }
    }
    else
    {
      // warp w owns the words [w * wpw, (w + 1) * wpw); candidates go through a 64-entry queue so that
      // all 32 lanes run an attempt at the same time
      const uint32_t wpw = NW / nwarps;
      uint32_t qn = 0;
      for (uint32_t r = 0; r < wpw; r += 32)
      {
        const uint32_t wi = wid * wpw + r + lane;
        uint32_t c = (r + lane < wpw) ? s_cand[wi] : 0;
        while (__any_sync(0xffffffffu, c != 0))
        {
          const bool has = c != 0;
          const uint32_t b = __ballot_sync(0xffffffffu, has);
          if (has)
          {
            const uint32_t k = __ffs(c) - 1;
            c &= c - 1;
            my_queue[qn + __popc(b & ((1u << lane) - 1))] = static_cast<uint16_t>(wi * 32 + k);
          }
          qn += __popc(b);
          __syncwarp();
          if (qn >= 32)
          {
            qn -= 32;
            const uint32_t off = my_queue[qn + lane];
            if (attempt<HAS_META>(t, P, T, tile_base + off))
            {
              atomicOr(&s_succ[off >> 5], 1u << (off & 31));
              any_succ = true;
            }
            __syncwarp();
          }
        }
      }
      if (lane < qn)
      {
        const uint32_t off = my_queue[lane];
        if (attempt<HAS_META>(t, P, T, tile_base + off))
        {
          atomicOr(&s_succ[off >> 5], 1u << (off & 31));
          any_succ = true;
        }
      }
    }
    // ---------------- phase C: successes -> lines
    const int tile_has_succ = __syncthreads_or(any_succ);
    // "position after the last newline so far" (exclusive over words), -1 = no line start yet in this tile
    constexpr uint32_t per = NW / THREADS; // 1 or 2 words per thread
    int local_last = -1;
    int own[per];
#pragma unroll
    for (uint32_t i = 0; i < per; ++i)
    {
      const uint32_t wi = threadIdx.x * per + i;
      const uint32_t w = s_nl[wi];
      own[i] = local_last;
      if (w != 0)
        local_last = static_cast<int>(wi * 32 + 32 - __clz(w));
    }
    int incl = warp_incl_max(local_last);
    if (lane == 31)
      s_warp_max[wid] = incl;
    __syncthreads();
    if (wid == 0)
    {
      int v = lane < nwarps ? s_warp_max[lane] : -1;
      int sc = warp_incl_max(v);
      int ex = __shfl_up_sync(0xffffffffu, sc, 1);
      if (lane == 0)
        ex = -1;
      if (lane < nwarps)
        s_warp_max[lane] = ex;
      if (lane == 31)
        s_last_start = sc;
    }
    __syncthreads();
    int before = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0)
      before = -1;
    before = max(before, s_warp_max[wid]);
    // the tile starts a line if it is the head of the buffer or follows a newline
    const int vstart = (tile_base == 0 || __ldg(buf + tile_base - 1) == '\n') ? 0 : -1;
    before = max(before, vstart);
    int last_start = max(s_last_start, vstart);
    if (tile_has_succ)
    {
#pragma unroll
      for (uint32_t i = 0; i < per; ++i)
      {
        const uint32_t wi = threadIdx.x * per + i;
        uint32_t s = s_succ[wi];
        const uint32_t w = s_nl[wi];
        const int prev = max(own[i], before);
        while (s != 0)
        {
          const uint32_t bit = __ffs(s) - 1;
          s &= s - 1;
          const uint32_t below = w & ((1u << bit) - 1);
          const int start = below != 0 ? static_cast<int>(wi * 32 + 32 - __clz(below)) : prev;
          if (start >= 0)
            atomicOr(&s_line[start >> 5], 1u << (start & 31));
        }
      }
      __syncthreads();
      for (uint32_t wi = threadIdx.x; wi < NW; wi += blockDim.x)
        my_lines += __popc(s_line[wi]);
    }
    // ---------------- the line that runs past the end of the tile
    const uint64_t tile_end = tile_base + ANY_TILE;
    bool open = tile_end < n && last_start >= 0 && last_start < static_cast<int>(ANY_TILE);
    if (open && tile_has_succ)
      open = ((s_line[last_start >> 5] >> (last_start & 31)) & 1u) == 0;
    if (open) // uniform across the CTA
    {
      uint64_t pos0 = tile_end;
      for (;;)
      {
        __syncthreads();
        if (threadIdx.x == 0)
          s_first_nl = 0xffffffffu;
        __syncthreads();
        const uint64_t base = pos0 + static_cast<uint64_t>(threadIdx.x) * 16;
        uint32_t cm = 0, nl = 0;
        if (base < n)
        {
          Window W;
          if (load_window(buf, n, base, W))
            cm = chunk_cand_fast(W, t, P, T, base);
          else
            cm = chunk_cand_generic(t, P, T, base);
          nl = newline_mask16(W);
          if (base + 16 > n)
            nl &= (1u << (n - base)) - 1;
          if (nl != 0)
            atomicMin(&s_first_nl, threadIdx.x * 16 + __ffs(nl) - 1);
        }
        __syncthreads();
        const uint32_t first_nl = s_first_nl;
        bool hit = false;
        while (cm != 0 && !hit)
        {
          const uint32_t k = __ffs(cm) - 1;
          cm &= cm - 1;
          if (threadIdx.x * 16 + k < first_nl)
            hit = attempt<HAS_META>(t, P, T, base + k);
        }
        if (__syncthreads_or(hit))
        {
          if (threadIdx.x == 0)
            ++my_lines;
          break;
        }
        pos0 += static_cast<uint64_t>(blockDim.x) * 16;
        if (first_nl != 0xffffffffu || pos0 >= n)
          break;
      }
    }
    __syncthreads();
  }
  // ---------------- totals: one atomic pair per CTA
  unsigned long long a = my_lines, b = my_newlines;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
  {
    a += __shfl_down_sync(0xffffffffu, a, d);
    b += __shfl_down_sync(0xffffffffu, b, d);
  }
  if (lane == 0)
  {
    atomicAdd(&s_red[0], a);
    atomicAdd(&s_red[1], b);
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    atomicAdd(&totals[0], s_red[0]);
    atomicAdd(&totals[1], s_red[1]);
  }
}

static size_t any_smem_bytes(const DevPattern& P, bool stage)
{
  return 256 + UGX_HASH + UGX_BTAP + 5 * (ANY_TILE / 8) + 32 * 64 * 2 + (stage ? ((P.table_bytes + 15) / 16) * 16 : 0);
}

bool count_lines_any_eligible(const DevPattern& P)
{
  return P.lbk == 0 && (P.flags & UGX_OPT_W) == 0 && P.adv != UGX_ADV_NONE;
}

cudaError_t launch_count_lines_any(const DevPattern& P, const uint8_t* buf, uint64_t n, unsigned long long* totals,
                                   int sm_count, cudaStream_t st)
{
  const uint64_t ntiles = (n + ANY_TILE - 1) / ANY_TILE;
  const bool stage = P.has_meta == 0 && any_smem_bytes(P, true) <= static_cast<size_t>(UGX_MAX_DYN_SMEM);
  const size_t smem = any_smem_bytes(P, stage);
  // big tables leave room for one CTA per SM: make it a full 1024-thread CTA; otherwise 512-thread CTAs
  const int threads = (!P.has_meta && smem > 100 * 1024) ? 1024 : 512;
  int per_sm = static_cast<int>((226 * 1024) / (smem + 1024));
  if (per_sm > 2048 / threads)
    per_sm = 2048 / threads;
  if (per_sm < 1)
    per_sm = 1;
  uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;
  if (g > ntiles)
    g = ntiles;
  if (g == 0)
    g = 1;
  cudaError_t e = cudaMemsetAsync(totals, 0, 2 * sizeof(unsigned long long), st);
  if (e != cudaSuccess)
    return e;
  const uint32_t st_flag = stage ? 1u : 0u;
#define UGX_LAUNCH_ANY(META, THR)                                                                                         \
  do                                                                                                                      \
  {                                                                                                                       \
    e = cudaFuncSetAttribute(count_lines_any_kernel<META, THR>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                             UGX_MAX_DYN_SMEM);                                                                     \
    if (e != cudaSuccess)                                                                                                 \
      return e;                                                                                                           \
    count_lines_any_kernel<META, THR><<<static_cast<int>(g), THR, smem, st>>>(P, buf, n, ntiles, st_flag, totals);         \
  } while (0)
  if (P.has_meta)
    UGX_LAUNCH_ANY(true, 512);
  else if (threads == 1024)
    UGX_LAUNCH_ANY(false, 1024);
  else
    UGX_LAUNCH_ANY(false, 512);
#undef UGX_LAUNCH_ANY
  return cudaGetLastError();
}

} // namespace ugx
