// scan_kernels.hpp — launch interface of the scan kernels (host side, CUDA runtime types only)
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ugrep_b200.h"

namespace ugx {

struct DevPattern;

constexpr int SCAN_THREADS = 256;                      // threads per CTA (512 for big staged tables: scan_threads())
// Dynamic shared memory every kernel opts in to (the 227 KiB per-CTA limit minus room for static shared memory).  The
// attribute is set to this one value for every pattern: it is process-global per kernel, and scanners on different
// host threads may launch the same kernel with different tables.
constexpr int UGX_MAX_DYN_SMEM = 227 * 1024 - 4096;
constexpr int SCAN_STRIP = 64;                         // bytes per thread per tile
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_STRIP;   // bytes per CTA iteration (16 KiB)
constexpr uint32_t SCAN_LINE_CAP = 1024;                // line starts per tile the dense line list holds
constexpr uint32_t SCAN_MAX_SMEM_TABLE = 198 * 1024;   // largest transition table staged in shared memory
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

// CTA size / tile size of the line-scan and records kernels for this pattern (ntiles = ceil(n / scan_tile_bytes))
int scan_threads(const DevPattern& P);
uint32_t scan_tile_bytes(const DevPattern& P);
cudaError_t launch_scan_lines(const DevPattern& P, const ScanArgs& a, int mode, bool emit, int sm_count, cudaStream_t st);
// `ugrep -c` for patterns without look-back: one position-parallel kernel, totals[0] = matching lines, [1] = newlines
bool count_lines_any_eligible(const DevPattern& P);
cudaError_t launch_count_lines_any(const DevPattern& P, const uint8_t* buf, uint64_t n, unsigned long long* totals,
                                   int sm_count, cudaStream_t st);
// ---- streaming `ugrep -c` (stream_count.cu) ----
#ifndef UGX_STREAM_THREADS
#define UGX_STREAM_THREADS 256
#endif
#ifndef UGX_SC_REGION
#define UGX_SC_REGION 16384
#endif
constexpr int STREAM_THREADS = UGX_STREAM_THREADS;     // threads per CTA
constexpr uint32_t SC_REGION = UGX_SC_REGION;                  // bytes per region (unit of the ticket counter and of the summaries)
constexpr uint32_t STREAM_MAX_GRID = 4096;             // capacity of the per-CTA partials

struct StreamArgs {
  uint8_t* region_sum;            // [regions(buffer) + 16] one summary byte per region of the buffer
  uint64_t region_begin;          // this launch scans the regions [region_begin, region_end) of the buffer;
  uint64_t region_end;            //   the kernel's n = the bytes of the buffer that are valid (copied) so far
  unsigned long long* partials;   // [2 * STREAM_MAX_GRID] per-CTA {lines, newlines}
  unsigned long long* ticket;     // region ticket counter (zero between launches)
  unsigned int* done;             // finished-CTA counter (zero between launches)
  unsigned long long* totals;     // [2] {matching lines, newlines}
  uint32_t finalize;              // last launch of a buffer: chain the regions' head lines
  uint32_t accumulate;            // add to totals instead of overwriting them
  uint32_t stage_table;           // set by the launcher
  uint32_t use_h4;                // set by the launcher: stage the hashed-predictor term table
  uint32_t use_via;               // set by the launcher: stage the k-gram viability tables (viability.cuh)
  uint32_t no_cover;              // evaluate the candidate predicate even when DevPattern::covers allows skipping it (A/B)
};

bool count_lines_stream_eligible(const DevPattern& P);
uint64_t stream_regions(uint64_t n);
int stream_grid(uint64_t n, int sm_count, int per_sm, int threads);
cudaError_t launch_count_lines_stream(const DevPattern& P, const uint8_t* buf, uint64_t n, StreamArgs a, bool want_nl,
                                      int sm_count, cudaStream_t st);
// the literal specialisation (stream_literal.cu), taken by launch_count_lines_stream when eligible
bool count_lines_literal_eligible(const DevPattern& P);
cudaError_t launch_count_lines_literal(const DevPattern& P, const uint8_t* buf, uint64_t n, const StreamArgs& a, bool want_nl,
                                       int sm_count, cudaStream_t st);

// find loop with position-parallel attempts (match_lines.cu): counting for patterns without META / W / start-loop skip
bool match_lines_eligible(const DevPattern& P);
uint32_t match_lines_tile_bytes(const DevPattern& P);
cudaError_t launch_match_lines(const DevPattern& P, const ScanArgs& a, int mode, int sm_count, cudaStream_t st);

// single-pass records (records_kernel.cu): staging pass + reorder into input order
cudaError_t launch_scan_records(const DevPattern& P, const ScanArgs& a, uint64_t* tile_base, ugx_match* stage_out,
                                uint64_t stage_cap, unsigned long long* cursor, int sm_count, cudaStream_t st);
cudaError_t launch_reorder_records(const ugx_match* stage, ugx_match* out, const uint64_t* pm, const uint64_t* pn,
                                   const uint64_t* tile_base, uint64_t ntiles, const unsigned long long* totals,
                                   uint64_t base_line, int sm_count, cudaStream_t st);

// ---- span scan (span_scan.cu): `ugrep -c -o` / `ugrep -o -n -b` with every DFA attempt taken out of the find loop ----
constexpr uint64_t SPAN_TAIL_MAX = 65536;   // a last line up to this long is left to the final kernel's line-at-a-time form

struct SpanArgs {
  uint64_t* reg_matches;   // [regions] count pass: matches that start in the region; after the final kernel: exclusive prefix
  uint64_t* reg_newlines;  // [regions] likewise, newlines
  uint64_t* reg_emain;     // [regions] farthest end of a success that starts in spans 0..30 of the region
  uint64_t* reg_elast;     // [regions] ... in its last span (the next region's window)
  uint64_t* reg_v;         // [regions] validation point of the region's chain start (~0 = none needed)
  uint16_t* sel_bits;      // [ceil(n / 16) + 32] or nullptr: the selected match starts of every 16-byte chunk (records)
  ugx_match* out;          // emit pass: records in input order
  uint64_t out_cap;
  uint64_t base_offset, base_line;
  const uint64_t* tail;    // device: [0] = start of the last line (the spans scan [0, tail[0]))
  unsigned int* flags;     // device: bit 0 an attempt failed at the end of the buffer, bit 1 a match too long for the span tables
  uint32_t stage_table;    // set by the launcher
  uint32_t use_via;        // set by the launcher: the viability tables are staged in shared memory
  uint32_t no_cover;       // decide membership in the attempt set even when DevPattern::covers makes it moot (A/B)
};

bool span_scan_eligible(const DevPattern& P);
cudaError_t launch_last_line(const uint8_t* buf, uint64_t n, uint64_t* tail, cudaStream_t st);
cudaError_t launch_span_scan(const DevPattern& P, const uint8_t* buf, uint64_t n, SpanArgs a, bool emit, int sm_count,
                             cudaStream_t st);
cudaError_t launch_span_final(const DevPattern& P, const uint8_t* buf, uint64_t n, const SpanArgs& a, bool emit,
                              unsigned long long* totals, cudaStream_t st);

// reflex::nlcount (newline_count.cu)
cudaError_t launch_count_newlines(const uint8_t* buf, uint64_t n, unsigned long long* total, int sm_count, cudaStream_t st);

// ---- many files in one launch (batch_kernel.cu) ----
struct BatchArgs {
  const uint64_t* begins;        // [nfiles] device: offset of every file in the batch buffer (multiples of 16)
  const uint64_t* lens;          // [nfiles] device: its length
  const uint32_t* tile_file;     // [ntiles] device: the file a tile belongs to
  const uint32_t* tile_index;    // [ntiles] device: its index within that file
  uint64_t ntiles;
  unsigned long long* counts;    // [nfiles] device, zeroed by the caller
  uint32_t stage_table;          // set by the launcher
};
cudaError_t launch_scan_batch(const DevPattern& P, const uint8_t* buf, BatchArgs a, int mode, int sm_count, cudaStream_t st);

// reflex::isutf8 / NUL test (utf8_check.cu): flags bit 0 = not UTF-8 by the reference's rule, bit 1 = has a NUL
cudaError_t launch_utf8_check(const uint8_t* buf, uint64_t n, unsigned int* flags, int sm_count, cudaStream_t st);

cudaError_t launch_tile_prefix(uint64_t* tile_matches, uint64_t* tile_newlines, uint64_t ntiles,
                               unsigned long long* totals, cudaStream_t st);

} // namespace ugx
