// scan_kernels.cu — sm_100a scan kernels of the buffer-scan path.
//
// Work decomposition (DESIGN.md "line-parallel scan"): no DFA in scope has a transition
// on '\n' and the look-back set never contains '\n' (checked at pattern upload), so every
// line is an independent unit of Matcher::match(FIND) (lib/matcher.cpp:42-750).  The buffer
// is cut into 64-byte strips; a thread owns every line that STARTS in its strip and runs the
// reference's find loop on it: prefilter candidate -> look-back -> DFA -> resume.
//
// Kernels here:
//   scan_lines_kernel    generic exact line scan (all prefilter families, look-back, META, W)
//   tile_prefix_kernel   exclusive prefix of per-tile match / newline counts (+ totals)
// Records (ugrep -o) come from a second run of scan_lines_kernel that knows every strip's
// output offset: deterministic input order, no atomics in the ordering.
#include "device_pattern.cuh"
#include "scan_kernels.hpp"

namespace ugx {

namespace {

constexpr int LINE_DONE = 2;

__device__ __forceinline__ bool advance_to(const Text& t, const DevPattern& P, const Tables& T, Cursor& m,
                                           uint64_t loc, uint64_t last)
{
  // first candidate in [loc, last]; `last` is the line's '\n' (or the final byte of the buffer)
  for (uint64_t k = loc; k <= last && k < t.end; ++k)
  {
    if (cand(t, P, T, k))
    {
      set_current(t, m, k);
      return true;
    }
  }
  return false;
}

// one anchored attempt over the dense table (patterns without META edges); lib/matcher.cpp:125-150, 446-546
__device__ __forceinline__ int run_dfa_table(const Text& t, const DevPattern& P, const Tables& T, Cursor& m, uint32_t& retry)
{
  const bool W = (P.flags & UGX_OPT_W) != 0;
  m.cap = 0;
  if (W && !at_wb(t, P, m))
    return 0;
  uint32_t state = 0;
  for (;;)
  {
    uint32_t acc = __ldg(P.accept + state);
    if ((acc & 0x7fffffffu) != 0 && (!W || at_we(t, P, peek_ch(t, m), m.pos)))
    {
      m.cap = acc & 0x7fffffffu;
      m.cur = m.pos;
    }
    if (acc & 0x80000000u) // state without outgoing edges: HALT before reading
      break;
    if (m.pos >= t.end)
      break;
    uint32_t ch = t.raw(m.pos++);
    uint32_t nxt = T.next[state * P.ncls + T.cls[ch]];
    if (nxt == D_DEAD)
      break;
    if (nxt == 0 && m.cap == 0) // back at the start state without an accept, lib/matcher.cpp:504-527
    {
      if (m.cur + 1 == m.pos)
      {
        ++m.cur;
        if (retry > 0)
          --retry;
      }
      else
      {
        while (m.cur + 1 < m.pos && !bit256(P.fst, t.raw(m.cur + 1)))
        {
          ++m.cur;
          if (retry > 0)
            --retry;
        }
      }
    }
    state = nxt;
  }
  return 0;
}

// one anchored attempt with the opcode interpreter (patterns with META edges); lib/matcher.cpp:94-546
__device__ int run_dfa_opc(const Text& t, const DevPattern& P, Cursor& m, uint32_t& retry)
{
  const bool W = (P.flags & UGX_OPT_W) != 0;
  const uint32_t* __restrict__ opc = P.opc;
  int ch = m.got;
  const bool bol = m.got == '\n';
  m.cap = 0;
  if (W && !at_wb(t, P, m))
    return 0;
  if (P.bol && !bol) // ^-anchored pattern away from a line start: the rest of this line cannot match
    return LINE_DONE;
  uint32_t pc = 0;
  uint32_t back = D_NONE;
  uint64_t bpos = 0;
  for (;;)
  {
    uint32_t op = __ldg(opc + pc);
    uint32_t jump;
    if (!d_op_is_goto(op))
    {
      if ((op >> 24) == 0xfe)
      {
        if (!W || at_we(t, P, peek_ch(t, m), m.pos))
        {
          m.cap = op & 0xffffff;
          m.cur = m.pos;
        }
        ++pc;
        continue;
      }
      if (ch == D_EOF)
        break;
      ch = get_ch(t, m);
      int metas = 5;
      jump = D_NONE;
      for (;;)
      {
        if (jump == D_NONE || back == D_NONE)
        {
          if (!d_op_is_goto(op))
          {
            uint32_t code = op >> 24;
            if (code == 0xfe)
            {
              if (!W || at_we(t, P, ch, m.pos - 1))
              {
                m.cap = op & 0xffffff;
                m.cur = m.pos;
                if (ch != D_EOF)
                  --m.cur;
              }
            }
            else if (code != 0xff)
            {
              if (metas > 0 && jump == D_NONE && meta_holds(t, P, m, code, ch, bol))
              {
                --metas;
                jump = op & 0xffff;
                if (jump == D_IDX_LONG)
                  jump = __ldg(opc + ++pc) & 0xffffff;
              }
            }
            op = __ldg(opc + ++pc);
            continue;
          }
          else if (ch != D_EOF && op != D_OP_HALT)
          {
            if (jump == D_NONE)
              break;
            if (back == D_NONE)
            {
              back = pc;
              bpos = m.pos - m.txt - 1;
            }
          }
        }
        if (jump == D_NONE)
        {
          if (back != D_NONE && bpos + 1 == m.pos - m.txt)
          {
            pc = back;
            op = __ldg(opc + pc);
            back = D_NONE;
          }
          break;
        }
        if (back == pc)
          bpos = m.pos - m.txt - 1;
        pc = jump;
        op = __ldg(opc + pc);
        jump = D_NONE;
      }
      if (ch == D_EOF)
        break;
    }
    else
    {
      if (op == D_OP_HALT)
      {
        if (back != D_NONE)
        {
          m.pos = m.txt + bpos;
          pc = back;
          back = D_NONE;
          continue;
        }
        break;
      }
      if (ch == D_EOF)
        break;
      ch = get_ch(t, m);
      if (ch == D_EOF)
        break;
    }
    while (static_cast<uint32_t>(ch) < (op >> 24) || static_cast<uint32_t>(ch) > ((op >> 16) & 0xff))
      op = __ldg(opc + ++pc);
    jump = op & 0xffff;
    if (jump == 0)
    {
      if (m.cap == 0)
      {
        if (m.cur + 1 == m.pos)
        {
          ++m.cur;
          if (retry > 0)
            --retry;
        }
        else
        {
          while (m.cur + 1 < m.pos && !bit256(P.fst, t.raw(m.cur + 1)))
          {
            ++m.cur;
            if (retry > 0)
              --retry;
          }
        }
      }
    }
    else if (jump >= D_IDX_LONG)
    {
      if (jump == D_IDX_HALT)
      {
        if (back != D_NONE)
        {
          pc = back;
          m.pos = m.txt + bpos;
          back = D_NONE;
          continue;
        }
        break;
      }
      jump = __ldg(opc + pc + 1) & 0xffffff;
    }
    pc = jump;
  }
  return 0;
}

__device__ __forceinline__ uint32_t look_back(const Text& t, const DevPattern& P, Cursor& m, uint64_t floor_pos)
{
  // walk back over cbk_ bytes from cur-1 down to floor_pos; lib/matcher.cpp:54-70, 639-654
  uint32_t retry = 0;
  uint64_t s = m.cur;
  if (s > floor_pos)
  {
    uint64_t n = s - floor_pos;
    if (P.lbk != 0xffff && P.lbk < n)
      n = P.lbk;
    while (n-- > 0 && bit256(P.cbk, t.raw(s - 1)))
    {
      --s;
      ++retry;
    }
    m.cur -= retry;
    retry = retry > P.lbm ? retry - P.lbm : 0;
  }
  return retry;
}

// one Matcher::match(FIND) confined to the line whose '\n' (or last byte) is at `last`.
// returns the accept index, or 0 when the line has no further match
template <bool HAS_META>
__device__ uint32_t find_in_line(const Text& t, const DevPattern& P, const Tables& T, Cursor& m, uint64_t last)
{
  const bool W = (P.flags & UGX_OPT_W) != 0;
  uint32_t retry = 0;
  m.len = 0;
  m.txt = m.cur;
  if (!advance_to(t, P, T, m, m.cur, last))
    return 0;
  if (P.lbk > 0)
  {
    retry = look_back(t, P, m, m.txt);
  }
  else if (P.one)
  {
    uint64_t k = m.cur + P.len;
    int ch = k < t.end ? static_cast<int>(t.raw(k)) : D_EOF;
    if (!W || (at_wb(t, P, m) && (m.pos >= t.end || at_we(t, P, ch, k))))
    {
      m.txt = m.cur;
      m.len = P.len;
      set_current(t, m, k);
      return m.cap = 1;
    }
  }
  set_current(t, m, m.cur);
  for (;;)
  {
    m.txt = m.cur;
    int r = HAS_META ? run_dfa_opc(t, P, m, retry) : run_dfa_table(t, P, T, m, retry);
    if (r == LINE_DONE)
      return 0;
    if (m.cap == 0)
    {
      if (m.pos < t.end)
      {
        if (retry > 0)
        {
          --retry;
          set_current(t, m, m.cur + 1);
          continue;
        }
        if (m.cur < m.pos)
        {
          if (!advance_to(t, P, T, m, m.cur + 1, last))
            return 0;
          if (P.lbk > 0)
          {
            retry = look_back(t, P, m, m.txt + 1);
            set_current(t, m, m.cur);
            continue;
          }
          if (!P.one)
            continue;
          uint64_t k = m.cur + P.len;
          int ch = k < t.end ? static_cast<int>(t.raw(k)) : D_EOF;
          if (W && (!at_wb(t, P, m) || !(m.pos >= t.end || at_we(t, P, ch, k))))
            continue;
          m.txt = m.cur;
          m.len = P.len;
          set_current(t, m, k);
          return m.cap = 1;
        }
      }
      m.txt = m.cur;
    }
    m.len = static_cast<uint32_t>(m.cur - m.txt);
    if (m.len == 0)
    {
      m.pos = m.cur;
      if (m.pos >= t.end)
        return 0;
      if (m.cap != 0)
      {
        if (!advance_to(t, P, T, m, m.cur + 1, last))
          return 0;
        continue;
      }
      if (m.cur + 1 > last)
        return 0;
      set_current(t, m, m.cur + 1);
      continue;
    }
    set_current(t, m, m.cur);
    return m.cap;
  }
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane)
{
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= static_cast<uint32_t>(d))
      x += y;
  }
  return x - v;
}

// exclusive scan over the block; total returned through *total (all threads)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total)
{
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t ex = warp_excl_scan(v, lane);
  if (lane == 31)
    warp_sums[wid] = ex + v;
  __syncthreads();
  if (wid == 0)
  {
    uint32_t s = lane < nw ? warp_sums[lane] : 0;
    uint32_t e = warp_excl_scan(s, lane);
    if (lane < nw)
      warp_sums[lane] = e;
    if (lane == 31)
      warp_sums[32] = e + s;
  }
  __syncthreads();
  uint32_t r = ex + warp_sums[wid];
  *total = warp_sums[32];
  __syncthreads();
  return r;
}

} // namespace

// MODE 0: count matching lines (ugrep -c), 1: count matches (ugrep -c -o) / emit records (ugrep -o)
template <int MODE, bool EMIT, bool HAS_META>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_lines_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n, uint64_t ntiles,
                  uint32_t stage_table, uint64_t* __restrict__ tile_matches, uint64_t* __restrict__ tile_newlines,
                  uint32_t* __restrict__ strip_counts, ugx_match* __restrict__ out, uint64_t out_cap,
                  uint64_t base_offset, uint64_t base_line)
{
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t warp_sums[33];
  // ---- stage the tables: class map, predictor, bitap pairs, and the transition table when it fits ----
  uint8_t* s_cls = smem;
  uint8_t* s_pred = smem + 256;
  uint8_t* s_tap = s_pred + UGX_HASH;
  uint16_t* s_next = reinterpret_cast<uint16_t*>(s_tap + UGX_BTAP);
  for (uint32_t i = threadIdx.x; i < 256 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(s_cls)[i] = __ldg(reinterpret_cast<const uint32_t*>(P.cls) + i);
  for (uint32_t i = threadIdx.x; i < UGX_HASH / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_pred)[i] = __ldg(reinterpret_cast<const uint4*>(P.pred) + i);
  for (uint32_t i = threadIdx.x; i < UGX_BTAP / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_tap)[i] = __ldg(reinterpret_cast<const uint4*>(P.tap) + i);
  if (stage_table)
    for (uint32_t i = threadIdx.x; i < (P.table_bytes + 15) / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(s_next)[i] = __ldg(reinterpret_cast<const uint4*>(P.next) + i);
  __syncthreads();
  Tables T;
  T.cls = s_cls;
  T.pred = s_pred;
  T.tap = s_tap;
  T.next = stage_table ? s_next : P.next;
  Text t{buf, n};

  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const uint64_t s0 = tile * SCAN_TILE + static_cast<uint64_t>(threadIdx.x) * SCAN_STRIP;
    // ---- newline mask of my strip and the line starts in it ----
    uint64_t nl = 0;
    if (s0 < n)
    {
      if (s0 + SCAN_STRIP <= n)
      {
        const uint4* q = reinterpret_cast<const uint4*>(buf + s0);
#pragma unroll
        for (int j = 0; j < SCAN_STRIP / 16; ++j)
        {
          uint4 v = __ldg(q + j);
          uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
          {
            uint32_t e = __vcmpeq4(w[k], 0x0a0a0a0au); // 0xff per equal byte
            // gather the top bit of each byte into 4 bits
            uint32_t bits = ((e & 0x80u) >> 7) | ((e & 0x8000u) >> 14) | ((e & 0x800000u) >> 21) | ((e & 0x80000000u) >> 28);
            nl |= static_cast<uint64_t>(bits) << (j * 16 + k * 4);
          }
        }
      }
      else
      {
        for (uint32_t i = 0; s0 + i < n; ++i)
          if (__ldg(buf + s0 + i) == '\n')
            nl |= 1ull << i;
      }
    }
    uint64_t starts = nl << 1;
    if (s0 < n && (s0 == 0 || __ldg(buf + s0 - 1) == '\n'))
      starts |= 1ull;
    if (s0 + SCAN_STRIP > n && s0 < n)
      starts &= (n - s0 >= 64) ? ~0ull : ((1ull << (n - s0)) - 1); // a line cannot start at or past the end
    const uint32_t my_nl = __popcll(nl);

    uint32_t strip_total = 0;
    uint64_t out_pos = 0;
    uint64_t line_no = 0;
    if (EMIT)
    {
      uint32_t cnt = s0 < n ? strip_counts[s0 / SCAN_STRIP] : 0;
      uint32_t tot;
      uint32_t ex = block_excl_scan(cnt, warp_sums, &tot);
      out_pos = tile_matches[tile] + ex;
      uint32_t nlex = block_excl_scan(my_nl, warp_sums, &tot);
      line_no = tile_newlines[tile] + nlex + 1 + base_line;
    }

    // ---- every line that starts in my strip ----
    uint64_t rest = starts;
    while (rest != 0)
    {
      const uint32_t bit = __ffsll(static_cast<long long>(rest)) - 1;
      rest &= rest - 1;
      const uint64_t L = s0 + bit;
      // end of the line: its '\n', or the last byte of the buffer
      uint64_t last;
      {
        const uint64_t after = nl >> bit; // newlines at or after L inside the strip
        if (after != 0)
          last = L + (__ffsll(static_cast<long long>(after)) - 1);
        else
        {
          uint64_t p = s0 + SCAN_STRIP;
          while (p < n && __ldg(buf + p) != '\n')
            ++p;
          last = p < n ? p : n - 1;
        }
      }
      uint64_t this_line = 0;
      if (EMIT)
        this_line = line_no + __popcll(nl & ((1ull << bit) - 1));
      Cursor m;
      set_current(t, m, L);
      for (;;)
      {
        uint32_t cap = find_in_line<HAS_META>(t, P, T, m, last);
        if (cap == 0)
          break;
        if (EMIT)
        {
          uint64_t idx = out_pos + strip_total;
          if (idx < out_cap)
          {
            ugx_match r;
            r.line = this_line;
            r.offset = m.txt + base_offset;
            r.len = m.len;
            r.cap = cap;
            out[idx] = r;
          }
        }
        ++strip_total;
        if (MODE == 0)
          break;
      }
    }

    if (!EMIT)
    {
      if (strip_counts != nullptr && s0 < n)
        strip_counts[s0 / SCAN_STRIP] = strip_total;
      // tile totals
      uint32_t tm = strip_total, tn = my_nl;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
      {
        tm += __shfl_down_sync(0xffffffffu, tm, d);
        tn += __shfl_down_sync(0xffffffffu, tn, d);
      }
      __shared__ uint32_t red_m[SCAN_THREADS / 32], red_n[SCAN_THREADS / 32];
      if ((threadIdx.x & 31) == 0)
      {
        red_m[threadIdx.x >> 5] = tm;
        red_n[threadIdx.x >> 5] = tn;
      }
      __syncthreads();
      if (threadIdx.x == 0)
      {
        uint64_t a = 0, b = 0;
        for (uint32_t w = 0; w < blockDim.x / 32; ++w)
        {
          a += red_m[w];
          b += red_n[w];
        }
        tile_matches[tile] = a;
        tile_newlines[tile] = b;
      }
      __syncthreads();
    }
  }
}

// exclusive prefix over the tiles (in place) + grand totals; one block
__global__ void __launch_bounds__(1024)
tile_prefix_kernel(uint64_t* __restrict__ tile_matches, uint64_t* __restrict__ tile_newlines, uint64_t ntiles,
                   unsigned long long* __restrict__ totals)
{
  __shared__ uint64_t sm[1024], sn[1024];
  const uint64_t per = (ntiles + blockDim.x - 1) / blockDim.x;
  const uint64_t lo = threadIdx.x * per;
  const uint64_t hi = lo + per < ntiles ? lo + per : ntiles;
  uint64_t a = 0, b = 0;
  for (uint64_t i = lo; i < hi; ++i)
  {
    a += tile_matches[i];
    b += tile_newlines[i];
  }
  sm[threadIdx.x] = a;
  sn[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint64_t ra = 0, rb = 0;
    for (uint32_t i = 0; i < blockDim.x; ++i)
    {
      uint64_t x = sm[i], y = sn[i];
      sm[i] = ra;
      sn[i] = rb;
      ra += x;
      rb += y;
    }
    totals[0] = ra;
    totals[1] = rb;
  }
  __syncthreads();
  a = sm[threadIdx.x];
  b = sn[threadIdx.x];
  for (uint64_t i = lo; i < hi; ++i)
  {
    uint64_t x = tile_matches[i], y = tile_newlines[i];
    tile_matches[i] = a;
    tile_newlines[i] = b;
    a += x;
    b += y;
  }
}

// ---- launchers ----

static size_t scan_smem_bytes(const DevPattern& P, bool stage)
{
  return 256 + UGX_HASH + UGX_BTAP + (stage ? ((P.table_bytes + 15) / 16) * 16 : 0);
}

template <int MODE, bool EMIT, bool HAS_META>
static cudaError_t launch_one(const DevPattern& P, const ScanArgs& a, bool stage, int grid, size_t smem, cudaStream_t st)
{
  auto kern = scan_lines_kernel<MODE, EMIT, HAS_META>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess)
    return e;
  kern<<<grid, SCAN_THREADS, smem, st>>>(P, a.buf, a.n, a.ntiles, stage ? 1u : 0u, a.tile_matches, a.tile_newlines,
                                         a.strip_counts, a.out, a.out_cap, a.base_offset, a.base_line);
  return cudaGetLastError();
}

cudaError_t launch_scan_lines(const DevPattern& P, const ScanArgs& a, int mode, bool emit, int sm_count, cudaStream_t st)
{
  const bool stage = P.has_meta == 0 && P.table_bytes <= SCAN_MAX_SMEM_TABLE;
  const size_t smem = scan_smem_bytes(P, stage);
  // persistent grid: as many CTAs as fit per SM, times the SM count, capped by the number of tiles
  int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
  if (per_sm > 2048 / SCAN_THREADS)
    per_sm = 2048 / SCAN_THREADS;
  if (per_sm < 1)
    per_sm = 1;
  uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;
  if (g > a.ntiles)
    g = a.ntiles;
  if (g == 0)
    g = 1;
  const int grid = static_cast<int>(g);
  const bool meta = P.has_meta != 0;
  if (mode == 0)
    return meta ? launch_one<0, false, true>(P, a, stage, grid, smem, st) : launch_one<0, false, false>(P, a, stage, grid, smem, st);
  if (!emit)
    return meta ? launch_one<1, false, true>(P, a, stage, grid, smem, st) : launch_one<1, false, false>(P, a, stage, grid, smem, st);
  return meta ? launch_one<1, true, true>(P, a, stage, grid, smem, st) : launch_one<1, true, false>(P, a, stage, grid, smem, st);
}

cudaError_t launch_tile_prefix(uint64_t* tile_matches, uint64_t* tile_newlines, uint64_t ntiles,
                               unsigned long long* totals, cudaStream_t st)
{
  tile_prefix_kernel<<<1, 1024, 0, st>>>(tile_matches, tile_newlines, ntiles, totals);
  return cudaGetLastError();
}

} // namespace ugx
