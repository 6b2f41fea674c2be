// batch_kernel.cu — scan_batch_kernel: MANY files in ONE launch.  The reference hands every file to
// a worker thread through its job queues (GrepMaster::submit / GrepWorker::execute, src/ugrep.cpp:4295-4432); one
// kernel launch per file would be hopeless for `ugrep -r` over a source tree, so the files are packed back to back
// (each starting on a 16-byte boundary) into one buffer with a table of {begin, length}, and one persistent grid walks
// all their 16 KiB tiles.  Every file keeps its own end of buffer — the prefilters' end-of-buffer rules, the
// unterminated last line and "no match crosses a file" are exactly those of a scan of the file alone: a tile's
// Text{} is the file, not the batch — and its own counter: what `ugrep -c` / `ugrep -c -o` prints for it
// (src/ugrep.cpp:10536-10586), summed per file with one atomic per tile.
//
// The tile body is the line-at-a-time form of scan_kernels.cu: phase A (position-parallel prefilter) writes the tile's
// candidate / newline bitmaps, then a thread runs the reference's find loop (find_in_line) on the lines that start in
// its 64-byte strip.
#include "device_pattern.cuh"
#include "line_match.cuh"
#include "ptx.cuh"
#include "scan_kernels.hpp"
#include "tile_phase_a.cuh"

namespace ugx {

// MODE 0: lines with a match per file, 1: matches per file
template <int MODE, bool HAS_META>
__global__ void __launch_bounds__(SCAN_THREADS, 4)
scan_batch_kernel(const __grid_constant__ DevPattern P, const uint8_t* __restrict__ buf, BatchArgs a)
{
  constexpr uint32_t TILE = SCAN_TILE;
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t red[SCAN_THREADS / 32];
  uint8_t* s_cls = smem;
  uint8_t* s_pred = smem + 256;
  uint8_t* s_tap = s_pred + UGX_HASH;
  uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_tap + UGX_BTAP);
  uint32_t* s_nl = s_cand + TILE / 32;
  uint16_t* s_next = reinterpret_cast<uint16_t*>(s_nl + TILE / 32);
  stage_tables_bulk(&s_bar, s_cls, P.cls, s_pred, P.pred, s_tap, P.tap, s_next, P.next,
                    a.stage_table ? ((P.table_bytes + 15) / 16) * 16 : 0);
  __syncthreads();
  Tables T;
  T.cls = s_cls;
  T.pred = s_pred;
  T.tap = s_tap;
  T.next = a.stage_table ? s_next : P.next;
  for (uint64_t g = blockIdx.x; g < a.ntiles; g += gridDim.x)
  {
    const uint32_t f = a.tile_file[g];
    const uint64_t begin = a.begins[f], len = a.lens[f];
    const uint64_t tile_base = static_cast<uint64_t>(a.tile_index[g]) * TILE; // within the file
    const Text t{buf + begin, len};
    const uint64_t s0 = tile_base + static_cast<uint64_t>(threadIdx.x) * SCAN_STRIP;
    tile_phase_a<TILE / 16 / SCAN_THREADS>(t, P, T, tile_base, reinterpret_cast<uint16_t*>(s_cand),
                                           reinterpret_cast<uint16_t*>(s_nl));
    __syncthreads();
    const CandMap cm{s_cand, tile_base, TILE};
    const uint64_t nl = (static_cast<uint64_t>(s_nl[2 * threadIdx.x + 1]) << 32) | s_nl[2 * threadIdx.x];
    uint64_t starts = nl << 1;
    if (s0 < len && (s0 == 0 || t.raw(s0 - 1) == '\n'))
      starts |= 1ull;
    if (s0 >= len)
      starts = 0;
    else if (len - s0 < 64)
      starts &= (1ull << (len - s0)) - 1;
    uint32_t mine = 0;
    uint64_t rest = starts;
    while (rest != 0)
    {
      const uint32_t bit = __ffsll(static_cast<long long>(rest)) - 1;
      rest &= rest - 1;
      const uint64_t L = s0 + bit;
      uint64_t last;
      const uint64_t after = nl >> bit;
      if (after != 0)
        last = L + (__ffsll(static_cast<long long>(after)) - 1);
      else
      {
        uint64_t p = s0 + SCAN_STRIP;
        while (p < len && t.raw(p) != '\n')
          ++p;
        last = p < len ? p : len - 1;
      }
      Cursor m;
      set_current(t, m, L);
      for (;;)
      {
        if (find_in_line<HAS_META>(t, P, T, cm, m, last) == 0)
          break;
        ++mine;
        if (MODE == 0)
          break;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
      mine += __shfl_down_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0)
      red[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      unsigned long long sum = 0;
      for (uint32_t w = 0; w < SCAN_THREADS / 32; ++w)
        sum += red[w];
      if (sum != 0)
        atomicAdd(a.counts + f, sum);
    }
    __syncthreads(); // the bitmaps are rewritten by the next tile
  }
}

cudaError_t launch_scan_batch(const DevPattern& P, const uint8_t* buf, BatchArgs a, int mode, int sm_count, cudaStream_t st)
{
  const bool stage = P.has_meta == 0 && P.table_bytes <= SCAN_MAX_SMEM_TABLE;
  const size_t smem = 256 + UGX_HASH + UGX_BTAP + 2 * (SCAN_TILE / 8) + (stage ? ((P.table_bytes + 15) / 16) * 16 : 0);
  a.stage_table = stage ? 1u : 0u;
  int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
  if (per_sm > 8)
    per_sm = 8;
  if (per_sm < 1)
    per_sm = 1;
  uint64_t g = static_cast<uint64_t>(sm_count) * per_sm;
  if (g > a.ntiles)
    g = a.ntiles;
  if (g == 0)
    g = 1;
#define UGX_BATCH_GO(MODE, META)                                                                            \
  do                                                                                                        \
  {                                                                                                         \
    auto kern = scan_batch_kernel<MODE, META>;                                                              \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UGX_MAX_DYN_SMEM); \
    if (e != cudaSuccess)                                                                                   \
      return e;                                                                                             \
    kern<<<static_cast<int>(g), SCAN_THREADS, smem, st>>>(P, buf, a);                                       \
    return cudaGetLastError();                                                                              \
  } while (0)
  if (P.has_meta)
  {
    if (mode == 0)
      UGX_BATCH_GO(0, true);
    UGX_BATCH_GO(1, true);
  }
  if (mode == 0)
    UGX_BATCH_GO(0, false);
  UGX_BATCH_GO(1, false);
#undef UGX_BATCH_GO
}

} // namespace ugx
