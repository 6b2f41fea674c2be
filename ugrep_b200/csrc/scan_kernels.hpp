// scan_kernels.hpp — launch interface of the scan kernels (host side, CUDA runtime types only)
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ugrep_b200.h"

namespace ugx {

struct DevPattern;

constexpr int SCAN_THREADS = 256;                      // threads per CTA
constexpr int SCAN_STRIP = 64;                         // bytes per thread per tile
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_STRIP;   // bytes per CTA iteration (16 KiB)
constexpr uint32_t SCAN_MAX_SMEM_TABLE = 200 * 1024;   // largest transition table staged in shared memory
constexpr uint32_t ANY_TILE = 32768;                   // bytes per CTA iteration of count_lines_any_kernel

struct ScanArgs {
  const uint8_t* buf;
  uint64_t n;
  uint64_t ntiles;
  uint64_t* tile_matches;   // [ntiles]
  uint64_t* tile_newlines;  // [ntiles]
  uint32_t* strip_counts;   // [ceil(n / SCAN_STRIP)] or nullptr
  ugx_match* out;
  uint64_t out_cap;
  uint64_t base_offset;
  uint64_t base_line;
};

cudaError_t launch_scan_lines(const DevPattern& P, const ScanArgs& a, int mode, bool emit, int sm_count, cudaStream_t st);
// `ugrep -c` for patterns without look-back: one position-parallel kernel, totals[0] = matching lines, [1] = newlines
bool count_lines_any_eligible(const DevPattern& P);
cudaError_t launch_count_lines_any(const DevPattern& P, const uint8_t* buf, uint64_t n, unsigned long long* totals,
                                   int sm_count, cudaStream_t st);
cudaError_t launch_tile_prefix(uint64_t* tile_matches, uint64_t* tile_newlines, uint64_t ntiles,
                               unsigned long long* totals, cudaStream_t st);

} // namespace ugx
