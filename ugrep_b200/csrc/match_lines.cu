// match_lines.cu — match_lines_kernel: the find loop of Matcher::match(FIND) (lib/matcher.cpp:42-750) with its DFA
// attempts taken OUT of the per-line sequential loop, for patterns with or without look-back (`ugrep -c -o`, and
// `ugrep -c` of look-back patterns; config 5).
//
// Why this is exact.  For patterns without META edges, without option W, that are not a pure literal and whose DFA
// never returns to its start state (no start-loop skip, lib/matcher.cpp:504-527), one find() is:
//     k  = first prefilter candidate >= cursor                                   (advance_*)
//     r  = bytes of the look-back set directly before k, at most lbk_ and not before `floor`  (:54-70, 627-658)
//     try anchored attempts at p = k - r, then p + 1, ... while the retry budget max(r - lbm_, 0) lasts;
//     a failed attempt with no budget left continues with the first candidate >= p + 1 and floor = p + 1;
//     the first attempt that accepts is the match, the cursor moves to its end.
// The attempt at p is a function of the text alone (longest accept of the DFA anchored at p), every attempted
// position is a candidate or a look-back byte, and within one find() the attempted positions only increase.  So:
//
//   phase A  prefilter + newline bitmaps of the tile (tile_phase_a.cuh), and a bitmap of the look-back bytes;
//   phase C  D(p) = length of the anchored match at p, for EVERY p that is a candidate or a look-back byte —
//            position-parallel;
//   phase D  per line, the sequential rules above on bitmaps and D(): no text access, no DFA.
//
// Lines that leave the tile, attempts that run into the end of the buffer and matches longer than 65534 bytes
// take the line-at-a-time form (find_in_line) instead (D() is 8 bits wide: matches of 255 bytes or more too).
#include "block_scan.cuh"
#include "device_pattern.cuh"
#include "line_match.cuh"
#include "ptx.cuh"
#include "scan_kernels.hpp"
#include "tile_phase_a.cuh"

namespace ugx {

namespace {

constexpr uint32_t ML_STRIP = 32;      // bytes per thread per tile (two 16-byte chunks in phase A)
constexpr uint32_t ML_FALLBACK = 0xffu; // D(p) is kept in 8 bits: longer matches take the line-at-a-time form

// first set bit of `bits` at a position in [from, limit] (tile offsets), or 0xffffffff
__device__ __forceinline__ uint32_t next_set(const uint32_t* bits, uint32_t from, uint32_t limit)
{
  if (from > limit)
    return 0xffffffffu;
  uint32_t wi = from >> 5;
  const uint32_t wl = limit >> 5;
  uint32_t word = bits[wi] & (0xffffffffu << (from & 31));
  while (word == 0 && wi < wl)
    word = bits[++wi];
  if (word == 0)
    return 0xffffffffu;
  const uint32_t hit = (wi << 5) + (__ffs(word) - 1);
  return hit <= limit ? hit : 0xffffffffu;
}

// number of consecutive set bits of `bits` directly before position k, at most maxr (k - maxr >= 0)
__device__ __forceinline__ uint32_t run_before(const uint32_t* bits, uint32_t k, uint32_t maxr)
{
  uint32_t r = 0;
  while (r < maxr)
  {
    const uint32_t q = k - r;            // looking at bits q-1, q-2, ...
    const uint32_t wi = (q - 1) >> 5;
    const uint32_t hi = (q - 1) & 31;    // highest bit of this word to look at
    // bits hi, hi-1, ..., 0 of the word, inverted: the first zero stops the run
    const uint32_t inv = ~bits[wi] << (31 - hi); // bit 31 = bit hi of the word
    const uint32_t ones = inv == 0 ? hi + 1 : __clz(inv);
    const uint32_t take = ones < maxr - r ? ones : maxr - r;
    r += take;
    if (take < hi + 1 || take == 0)
      break;
  }
  return r;
}

} // namespace

// MODE 0: lines with a match, 1: matches
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS >= 512 ? 2 : 3)
match_lines_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n, uint64_t ntiles,
                   uint32_t stage_table, uint64_t* __restrict__ tile_matches, uint64_t* __restrict__ tile_newlines)
{
  constexpr uint32_t TILE = THREADS * ML_STRIP;
  constexpr uint32_t NW = TILE / 32;           // bitmap words per tile
  constexpr uint32_t NWARPS = THREADS / 32;
  constexpr uint32_t LINE_CAP = THREADS * 2;   // dense line list
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t red_m[32], red_n[32];
  __shared__ __align__(8) uint64_t s_bar;
  uint8_t* s_cls = smem;
  uint8_t* s_pred = smem + 256;
  uint8_t* s_tap = s_pred + UGX_HASH;
  uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_tap + UGX_BTAP);
  uint32_t* s_nl = s_cand + NW;
  uint32_t* s_cbk = s_nl + NW;
  uint8_t* s_len = reinterpret_cast<uint8_t*>(s_cbk + NW);         // [TILE] D(p), valid where attempted
  uint16_t* s_lines = reinterpret_cast<uint16_t*>(s_len + TILE);   // [LINE_CAP]
  uint16_t* s_next = s_lines + LINE_CAP;
  stage_tables_bulk(&s_bar, s_cls, P.cls, s_pred, P.pred, s_tap, P.tap, s_next, P.next,
                    stage_table ? ((P.table_bytes + 15) / 16) * 16 : 0);
  Tables T;
  T.cls = s_cls;
  T.pred = s_pred;
  T.tap = s_tap;
  T.next = stage_table ? s_next : P.next;
  const Text t{buf, n};
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool has_lb = P.lbk > 0;

  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const uint64_t tile_base = tile * TILE;
    const uint32_t tile_len = n - tile_base < TILE ? static_cast<uint32_t>(n - tile_base) : TILE;
    // ---- phase A: candidate and newline bitmaps; look-back byte bitmap
    tile_phase_a<2>(t, P, T, tile_base, reinterpret_cast<uint16_t*>(s_cand), reinterpret_cast<uint16_t*>(s_nl));
    {
      const uint32_t off = threadIdx.x * ML_STRIP;
      uint32_t m = 0;
      if (has_lb)
        for (uint32_t i = 0; i < ML_STRIP && off + i < tile_len; ++i)
          if (bit256(P.cbk, t.raw(tile_base + off + i)))
            m |= 1u << i;
      s_cbk[threadIdx.x] = m;
    }
    __syncthreads();
    // ---- phase C: D(p) for every candidate or look-back position.  A lane owns one bitmap word (32 positions) and
    // runs its attempts back to back in ONE loop whose body is a single DFA transition: a lane that finishes an
    // attempt fetches its next position in the same iteration, so all lanes keep stepping together whatever the
    // lengths of their attempts.
    {
      uint32_t bits = s_cand[threadIdx.x] | s_cbk[threadIdx.x];
      const uint32_t base_off = threadIdx.x * 32;
      const uint8_t* __restrict__ tb = buf + tile_base;                         // positions are tile-relative, 32 bits
      const uint64_t left = n - tile_base;
      const uint32_t lim = left > 0xfffffff0ull ? 0xfffffff0u : static_cast<uint32_t>(left);
      const uint32_t first_acc = P.first_acc, ncls = P.ncls;
      bool active = false;
      uint32_t state = 0, off = 0, p = 0, acc_at = 0;
      for (;;)
      {
        if (!active)
        {
          if (bits == 0)
            break;
          const uint32_t k = __ffs(bits) - 1;
          bits &= bits - 1;
          off = base_off + k;
          p = off;
          acc_at = off;
          state = 0;
          active = true;
        }
        bool stop = false, hit_end = false;
        if (state >= first_acc) // (the start state does not accept: match_lines_eligible)
        {
          const uint32_t acc = __ldg(P.accept + state);
          if ((acc & 0x7fffffffu) != 0)
            acc_at = p;
          stop = (acc & 0x80000000u) != 0;
        }
        if (!stop)
        {
          if (p >= lim)
          {
            stop = true;
            hit_end = true;
          }
          else
          {
            const uint32_t ch = __ldg(tb + p);
            ++p;
            const uint32_t nxt = T.next[state * ncls + T.cls[ch]];
            if (nxt == D_DEAD)
              stop = true;
            else
              state = nxt;
          }
        }
        if (stop)
        {
          const uint32_t len = acc_at - off;
          s_len[off] = static_cast<uint8_t>((hit_end || len >= ML_FALLBACK) ? ML_FALLBACK : len);
          active = false;
        }
      }
    }
    // ---- line starts of the tile, compacted
    const uint32_t nlw = s_nl[threadIdx.x];
    uint32_t starts = nlw << 1;
    const uint64_t s0 = tile_base + static_cast<uint64_t>(threadIdx.x) * ML_STRIP;
    if (s0 < n && (s0 == 0 || (threadIdx.x > 0 ? (s_nl[threadIdx.x - 1] >> 31) != 0 : __ldg(buf + s0 - 1) == '\n')))
      starts |= 1u;
    if (s0 >= n)
      starts = 0;
    else if (n - s0 < ML_STRIP)
      starts &= (1u << (n - s0)) - 1;
    uint32_t total_lines;
    const uint32_t first_idx = block_excl_scan(static_cast<uint32_t>(__popc(starts)), warp_sums, &total_lines);
    // (block_excl_scan synchronises the CTA: D() of phase C is complete and visible from here on)
    uint32_t mine = 0; // matches (MODE 1) or matching lines (MODE 0) found by this thread
    const bool dense = total_lines <= LINE_CAP;
    if (dense)
    {
      uint32_t idx = first_idx;
      uint32_t todo = starts;
      while (todo != 0)
      {
        const uint32_t bit = __ffs(todo) - 1;
        todo &= todo - 1;
        s_lines[idx++] = static_cast<uint16_t>(threadIdx.x * ML_STRIP + bit);
      }
    }
    __syncthreads();
    // ---- phase D: the sequential rules of find() per line, on bitmaps and D()
    uint32_t li = threadIdx.x;
    uint32_t rest = starts;
    for (;;)
    {
      uint32_t off;
      if (dense)
      {
        if (li >= total_lines)
          break;
        off = s_lines[li];
        li += blockDim.x;
      }
      else
      {
        if (rest == 0)
          break;
        const uint32_t bit = __ffs(rest) - 1;
        rest &= rest - 1;
        off = threadIdx.x * ML_STRIP + bit;
      }
      // end of the line inside the tile?
      const uint32_t nlpos = next_set(s_nl, off, tile_len - 1);
      bool slow = nlpos == 0xffffffffu && tile_base + tile_len < n; // the line leaves the tile
      const uint32_t last = nlpos != 0xffffffffu ? nlpos : tile_len - 1; // tile offset of the line's last byte
      uint32_t found = 0;
      if (!slow)
      {
        uint32_t c = off;       // cursor
        for (;;)
        {
          // one find(): candidates from c
          uint32_t k = next_set(s_cand, c, last);
          if (k == 0xffffffffu)
            break;
          uint32_t floor_pos = c;
          uint32_t mlen = 0;
          for (;;)
          {
            uint32_t r = 0;
            if (has_lb && k > floor_pos)
            {
              uint32_t maxr = k - floor_pos;
              if (P.lbk != 0xffff && P.lbk < maxr)
                maxr = P.lbk;
              r = run_before(s_cbk, k, maxr);
            }
            uint32_t p = k - r;
            uint32_t retry = r > P.lbm ? r - P.lbm : 0;
            for (;;)
            {
              const uint32_t d = s_len[p];
              if (d == ML_FALLBACK)
              {
                slow = true;
                break;
              }
              if (d != 0)
              {
                mlen = d;
                break;
              }
              if (retry == 0)
                break;
              --retry;
              ++p;
            }
            if (slow || mlen != 0)
            {
              c = p; // match start (when mlen != 0)
              break;
            }
            // failed with no budget left: next candidate after p, look-back floor p + 1
            k = next_set(s_cand, p + 1, last);
            if (k == 0xffffffffu)
              break;
            floor_pos = p + 1;
          }
          if (slow || mlen == 0)
            break;
          ++found;
          if (MODE == 0)
            break;
          c += mlen; // the cursor moves to the end of the match
        }
      }
      if (slow)
      {
        // the line-at-a-time form for this line
        const uint64_t L = tile_base + off;
        uint64_t lastg;
        if (nlpos != 0xffffffffu)
          lastg = tile_base + nlpos;
        else
        {
          uint64_t q = tile_base + tile_len;
          while (q < n && __ldg(buf + q) != '\n')
            ++q;
          lastg = q < n ? q : n - 1;
        }
        const CandMap cm{s_cand, tile_base, tile_len};
        Cursor m;
        set_current(t, m, L);
        found = 0;
        for (;;)
        {
          if (find_in_line<false>(t, P, T, cm, m, lastg) == 0)
            break;
          ++found;
          if (MODE == 0)
            break;
        }
      }
      mine += found;
    }
    // ---- tile totals
    uint32_t tm = mine, tn = __popc(nlw);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
      tm += __shfl_down_sync(0xffffffffu, tm, d);
      tn += __shfl_down_sync(0xffffffffu, tn, d);
    }
    if (lane == 0)
    {
      red_m[wid] = tm;
      red_n[wid] = tn;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      uint64_t x = 0, y = 0;
      for (uint32_t w = 0; w < NWARPS; ++w)
      {
        x += red_m[w];
        y += red_n[w];
      }
      tile_matches[tile] = x;
      tile_newlines[tile] = y;
    }
    __syncthreads(); // bitmaps and D() are rewritten by the next tile
  }
}

bool match_lines_eligible(const DevPattern& P)
{
  return P.has_meta == 0 && (P.flags & UGX_OPT_W) == 0 && P.to_start == 0 && P.one == 0 && P.acc0 == 0 &&
         P.adv != UGX_ADV_NONE && P.table_bytes <= 150 * 1024;
}

static int match_lines_threads(const DevPattern& P) { return P.table_bytes > 24 * 1024 ? 512 : 256; }

uint32_t match_lines_tile_bytes(const DevPattern& P) { return static_cast<uint32_t>(match_lines_threads(P)) * ML_STRIP; }

template <int MODE, int THREADS>
static cudaError_t launch_ml(const DevPattern& P, const ScanArgs& a, int sm_count, cudaStream_t st)
{
  constexpr uint32_t TILE = THREADS * ML_STRIP;
  const size_t smem = 256 + UGX_HASH + UGX_BTAP + 3 * (TILE / 8) + TILE + 2 * (THREADS * 2) +
                      ((P.table_bytes + 15) / 16) * 16;
  auto kern = match_lines_kernel<MODE, THREADS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UGX_MAX_DYN_SMEM);
  if (e != cudaSuccess)
    return e;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
  if (e != cudaSuccess)
    return e;
  if (per_sm < 1)
    per_sm = 1;
  uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;
  if (g > a.ntiles)
    g = a.ntiles;
  if (g == 0)
    g = 1;
  kern<<<static_cast<int>(g), THREADS, smem, st>>>(P, a.buf, a.n, a.ntiles, 1u, a.tile_matches, a.tile_newlines);
  return cudaGetLastError();
}

cudaError_t launch_match_lines(const DevPattern& P, const ScanArgs& a, int mode, int sm_count, cudaStream_t st)
{
  if (match_lines_threads(P) == 512)
    return mode == 0 ? launch_ml<0, 512>(P, a, sm_count, st) : launch_ml<1, 512>(P, a, sm_count, st);
  return mode == 0 ? launch_ml<0, 256>(P, a, sm_count, st) : launch_ml<1, 256>(P, a, sm_count, st);
}

} // namespace ugx
