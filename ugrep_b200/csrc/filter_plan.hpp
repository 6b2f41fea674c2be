// filter_plan.hpp — the first-stage ("survivor") filter of the position-parallel scan, planned on the
// host from the compiled pattern's prefilter fields (plain C++: shared by host and device code).
//
// The reference decides whether byte k is a candidate with one of the advance_* routines
// (lib/matcher.cpp:957-3549).  On the GPU that predicate is evaluated in two stages:
//   stage 1 (every byte, branch-free):  a cheap SUPERSET test, described by a FilterPlan;
//   stage 2 (survivors only):           the exact predicate cand() of device_pattern.cuh.
// A plan never changes results: any position it drops fails the exact predicate too.
#pragma once

#include <cstdint>

namespace ugx {

enum FilterKind : uint32_t {
  FK_ALL = 0,      // no useful first stage: every position survives
  FK_NEVER = 1,    // the prefilter can never fire on an interior position (e.g. config 3)
  FK_ANCHOR2 = 2,  // literal prefix: two of its bytes compared at fixed offsets
  FK_LUT = 3       // one table lookup per byte position + up to 4 (offset, bit) terms
};

enum HashKind : uint32_t {
  HK_BYTE = 0, // L(p) = fb[text[p]]                         (256-byte set table built by the planner)
  HK_PAIR = 1, // L(p) = tap[(text[p] ^ text[p+1] << 6) & 2047]  (bitap pairs, Pattern::tap_)
  HK_H4 = 2    // L(p) = pmh[H(p)], H = 12-bit rolling hash of text[p-3..p] (Pattern::pmh_, steps j >= 3)
};

constexpr int FILTER_MAX_TERMS = 4;

// a position k survives iff for every term t: bit t_bit[t] of L(k + t_off[t]) is 0
struct FilterPlan {
  uint32_t kind;
  uint32_t hk;
  uint32_t nterms;
  uint32_t t_off[FILTER_MAX_TERMS]; // 0..8
  uint32_t t_bit[FILTER_MAX_TERMS]; // 0..7
  uint32_t p_lo, p_hi;              // lookups are needed for window positions [p_lo, p_hi), p_hi <= 24
  uint32_t a_off[2];                // FK_ANCHOR2: offsets (0..8) of the two compared bytes
  uint32_t a_chr[2];                // ... and the bytes, splatted over a word
  uint32_t est_pass_ppm;            // planner's estimate of the survivor rate (informational)
  uint8_t fb[256];                  // HK_BYTE table: bit b set = byte fails the set of term bit b
};

} // namespace ugx
