// records_kernel.cu — `ugrep -o [-n -b]`: match records (line, offset, length, accept) in input order from ONE
// run of the find loop per line (the two-pass form of scan_kernels.cu runs it twice).
//
//   scan_records_kernel   16 KiB tiles as in scan_lines_kernel: phase A bitmaps, then a thread runs the
//                         reference's find loop (Matcher::match(FIND), lib/matcher.cpp:42-750) on the lines that
//                         start in its 64-byte strip, in order, keeping up to REC_K matches of the strip in shared
//                         memory.  A block scan of the per-strip counts orders the tile's records; the tile takes
//                         its place in a staging buffer with one atomicAdd on a cursor (tiles land in completion
//                         order) and writes records with TILE-RELATIVE line numbers.  Strips with more than REC_K
//                         matches run their lines a second time, writing directly.
//   tile_prefix_kernel    (scan_kernels.cu) exclusive prefixes of the tiles' match / newline counts.
//   reorder_records_kernel moves every tile's records to its place in input order and adds the line-number base:
//                         what AbstractMatcher::lineno() returns (absmatcher.h:695-766) = 1 + newlines before.
// No ordering depends on an atomic: the cursor only assigns staging space.
#include "block_scan.cuh"
#include "device_pattern.cuh"
#include "ptx.cuh"
#include "line_match.cuh"
#include "scan_kernels.hpp"
#include "tile_phase_a.cuh"

namespace ugx {

namespace {

constexpr int REC_K = 4; // matches per strip kept in shared memory

struct StagedRec {
  uint32_t rel_off, len, cap, rel_line;
};

// end of the line that starts at tile offset `off`: the next newline in the tile's bitmap, else walk on in global
// memory; the last byte of the buffer if there is none
template <uint32_t TILE>
__device__ __forceinline__ uint64_t line_last(const uint32_t* s_nl, const uint8_t* __restrict__ buf, uint64_t n,
                                              uint64_t tile_base, uint32_t off)
{
  constexpr uint32_t NW = TILE / 32;
  uint32_t wi = off >> 5;
  uint32_t word = s_nl[wi] & (0xffffffffu << (off & 31));
  while (word == 0 && ++wi < NW)
    word = s_nl[wi];
  if (word != 0)
    return tile_base + (wi << 5) + (__ffs(word) - 1);
  uint64_t p = tile_base + TILE;
  while (p < n && __ldg(buf + p) != '\n')
    ++p;
  return p < n ? p : n - 1;
}

} // namespace

template <bool HAS_META, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS > 256 ? 2048 / THREADS : 8)
scan_records_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, uint64_t n, uint64_t ntiles,
                    uint32_t stage_table, uint64_t* __restrict__ tile_matches, uint64_t* __restrict__ tile_newlines,
                    uint64_t* __restrict__ tile_base_out, ugx_match* __restrict__ stage_out, uint64_t stage_cap,
                    unsigned long long* __restrict__ cursor, uint64_t base_offset)
{
  constexpr uint32_t TILE = THREADS * SCAN_STRIP;
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t warp_sums[33];
  __shared__ unsigned long long s_base;
  uint8_t* s_cls = smem;
  uint8_t* s_pred = smem + 256;
  uint8_t* s_tap = s_pred + UGX_HASH;
  uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_tap + UGX_BTAP);
  uint32_t* s_nl = s_cand + TILE / 32;
  StagedRec* s_rec = reinterpret_cast<StagedRec*>(s_nl + TILE / 32);
  uint16_t* s_next = reinterpret_cast<uint16_t*>(s_rec + THREADS * REC_K);
  // tables -> shared memory by bulk asynchronous copies (ptx.cuh)
  __shared__ __align__(8) uint64_t s_bar;
  stage_tables_bulk(&s_bar, s_cls, P.cls, s_pred, P.pred, s_tap, P.tap, s_next, P.next,
                    stage_table ? ((P.table_bytes + 15) / 16) * 16 : 0);
  __syncthreads();
  Tables T;
  T.cls = s_cls;
  T.pred = s_pred;
  T.tap = s_tap;
  T.next = stage_table ? s_next : P.next;
  const Text t{buf, n};

  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    const uint64_t tile_base = tile * TILE;
    const uint64_t s0 = tile_base + static_cast<uint64_t>(threadIdx.x) * SCAN_STRIP;
    tile_phase_a<TILE / 16 / THREADS>(t, P, T, tile_base, reinterpret_cast<uint16_t*>(s_cand),
                                                reinterpret_cast<uint16_t*>(s_nl));
    __syncthreads();
    const CandMap cm{s_cand, tile_base, TILE};
    const uint64_t nl = (static_cast<uint64_t>(s_nl[2 * threadIdx.x + 1]) << 32) | s_nl[2 * threadIdx.x];
    uint64_t starts = nl << 1;
    if (s0 < n && (s0 == 0 || __ldg(buf + s0 - 1) == '\n'))
      starts |= 1ull;
    if (s0 + SCAN_STRIP > n && s0 < n)
      starts &= (n - s0 >= 64) ? ~0ull : ((1ull << (n - s0)) - 1);
    const uint32_t my_nl = __popcll(nl);
    uint32_t tile_nl;
    const uint32_t nl_before = block_excl_scan(my_nl, warp_sums, &tile_nl); // newlines of the tile before my strip

    // ---- run 1: every line that starts in my strip, matches kept in shared memory
    uint32_t cnt = 0;
    bool spilled = false;
    StagedRec* mine = s_rec + threadIdx.x * REC_K;
    for (int run = 0; run < 2; ++run)
    {
      // run 1 (run == 1) happens only for strips that did not fit: it writes straight to the staging buffer
      uint64_t direct = 0;
      if (run == 1)
      {
        uint32_t tile_total;
        const uint32_t ex = block_excl_scan(cnt, warp_sums, &tile_total);
        if (threadIdx.x == 0)
        {
          s_base = atomicAdd(cursor, static_cast<unsigned long long>(tile_total));
          tile_matches[tile] = tile_total;
          tile_newlines[tile] = tile_nl;
          tile_base_out[tile] = s_base;
        }
        __syncthreads();
        const uint64_t base = s_base;
        const bool fits = base + tile_total <= stage_cap; // uniform over the CTA
        direct = base + ex;
        if (!fits)
          break;
        if (!spilled)
        {
          for (uint32_t i = 0; i < cnt; ++i)
          {
            ugx_match r;
            r.line = mine[i].rel_line;
            r.offset = tile_base + mine[i].rel_off + base_offset;
            r.len = mine[i].len;
            r.cap = mine[i].cap;
            stage_out[direct + i] = r;
          }
          break;
        }
      }
      uint32_t k = 0;
      uint64_t rest = starts;
      while (rest != 0)
      {
        const uint32_t bit = __ffsll(static_cast<long long>(rest)) - 1;
        rest &= rest - 1;
        const uint32_t off = threadIdx.x * SCAN_STRIP + bit;
        const uint64_t last = line_last<TILE>(s_nl, buf, n, tile_base, off);
        const uint32_t rel_line = nl_before + __popcll(nl & ((1ull << bit) - 1));
        Cursor m;
        set_current(t, m, tile_base + off);
        for (;;)
        {
          const uint32_t cap = find_in_line<HAS_META>(t, P, T, cm, m, last);
          if (cap == 0)
            break;
          if (run == 0)
          {
            const uint64_t rel = m.txt - tile_base;
            if (k < REC_K && rel <= 0xffffffffull)
              mine[k] = StagedRec{static_cast<uint32_t>(rel), m.len, cap, rel_line};
            else
              spilled = true;
          }
          else
          {
            ugx_match r;
            r.line = rel_line;
            r.offset = m.txt + base_offset;
            r.len = m.len;
            r.cap = cap;
            stage_out[direct + k] = r;
          }
          ++k;
        }
      }
      if (run == 0)
        cnt = k;
    }
    __syncthreads(); // the bitmaps and the record slots are rewritten by the next tile
  }
}

// one warp per tile: records of tile t go from the staging buffer to out[prefix_t ...], line numbers become absolute
__global__ void __launch_bounds__(256)
reorder_records_kernel(const ugx_match* __restrict__ stage, ugx_match* __restrict__ out, const uint64_t* __restrict__ pm,
                       const uint64_t* __restrict__ pn, const uint64_t* __restrict__ tile_base, uint64_t ntiles,
                       const unsigned long long* __restrict__ totals, uint64_t base_line)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t tl = warp; tl < ntiles; tl += nwarps)
  {
    const uint64_t first = pm[tl];
    const uint64_t count = (tl + 1 < ntiles ? pm[tl + 1] : totals[0]) - first;
    const ugx_match* src = stage + tile_base[tl];
    const uint64_t add = pn[tl] + 1 + base_line;
    for (uint64_t i = lane; i < count; i += 32)
    {
      ugx_match r = src[i];
      r.line += add;
      out[first + i] = r;
    }
  }
}

static size_t records_smem_bytes(const DevPattern& P, bool stage, int threads)
{
  return 256 + UGX_HASH + UGX_BTAP + 2 * (threads * SCAN_STRIP / 8) + threads * REC_K * sizeof(StagedRec) +
         (stage ? ((P.table_bytes + 15) / 16) * 16 : 0);
}

template <bool HAS_META, int THREADS>
static cudaError_t launch_records_one(const DevPattern& P, const ScanArgs& a, bool stage, int grid, size_t smem,
                                      uint64_t* tile_base, ugx_match* stage_out, uint64_t stage_cap,
                                      unsigned long long* cursor, cudaStream_t st)
{
  auto kern = scan_records_kernel<HAS_META, THREADS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UGX_MAX_DYN_SMEM);
  if (e != cudaSuccess)
    return e;
  kern<<<grid, THREADS, smem, st>>>(P, a.buf, a.n, a.ntiles, stage ? 1u : 0u, a.tile_matches, a.tile_newlines, tile_base,
                                    stage_out, stage_cap, cursor, a.base_offset);
  return cudaGetLastError();
}

cudaError_t launch_scan_records(const DevPattern& P, const ScanArgs& a, uint64_t* tile_base, ugx_match* stage_out,
                                uint64_t stage_cap, unsigned long long* cursor, int sm_count, cudaStream_t st)
{
  const int threads = scan_threads(P); // the tile size must agree with scan_tile_bytes(): the caller sized ntiles by it
  const bool stage = P.has_meta == 0 && P.table_bytes <= SCAN_MAX_SMEM_TABLE &&
                     records_smem_bytes(P, true, threads) <= static_cast<size_t>(UGX_MAX_DYN_SMEM);
  const size_t smem = records_smem_bytes(P, stage, threads);
  int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
  if (per_sm > 2048 / threads)
    per_sm = 2048 / threads;
  if (per_sm < 1)
    per_sm = 1;
  uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;
  if (g > a.ntiles)
    g = a.ntiles;
  if (g == 0)
    g = 1;
  cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess)
    return e;
  const int grid = static_cast<int>(g);
  if (P.has_meta)
    return launch_records_one<true, 256>(P, a, stage, grid, smem, tile_base, stage_out, stage_cap, cursor, st);
  if (threads != 256)
    return launch_records_one<false, 1024>(P, a, stage, grid, smem, tile_base, stage_out, stage_cap, cursor, st);
  return launch_records_one<false, 256>(P, a, stage, grid, smem, tile_base, stage_out, stage_cap, cursor, st);
}

cudaError_t launch_reorder_records(const ugx_match* stage, ugx_match* out, const uint64_t* pm, const uint64_t* pn,
                                   const uint64_t* tile_base, uint64_t ntiles, const unsigned long long* totals,
                                   uint64_t base_line, int sm_count, cudaStream_t st)
{
  uint64_t g = (ntiles + 7) / 8;
  if (g > static_cast<uint64_t>(sm_count) * 8)
    g = static_cast<uint64_t>(sm_count) * 8;
  if (g == 0)
    g = 1;
  reorder_records_kernel<<<static_cast<int>(g), 256, 0, st>>>(stage, out, pm, pn, tile_base, ntiles, totals, base_line);
  return cudaGetLastError();
}

} // namespace ugx
