// filter_plan.hpp — the first-stage ("survivor") filter of the position-parallel scans, planned on the
// host from the compiled pattern's prefilter fields (plain C++: shared by host and device code).
//
// The reference decides whether byte k is a candidate with one of the advance_* routines
// (lib/matcher.cpp:957-3549).  On the GPU that predicate is evaluated in two stages:
//   stage 1 (every byte, branch-free):  a cheap SUPERSET test described by a FilterPlan;
//   stage 2 (survivors only):           the exact predicate cand() of device_pattern.cuh.
// A plan never changes results: on interior positions (all bytes the routine reads exist) any position it drops
// fails the exact predicate too; near the end of the buffer the kernels do not use the plan.
#pragma once

#include <cstdint>

namespace ugx {

enum FilterKind : uint32_t {
  FK_ALL = 0,      // no useful first stage: every position survives
  FK_ANCHOR2 = 2,  // pure literal (Pattern::one_): two of its bytes compared at fixed offsets, SWAR
  FK_LUT = 3       // up to 3 byte-set tests at fixed offsets: position k survives iff for every term t
                   //   byte (k + t_off[t]) is in set t        (one shared-memory lookup per text byte)
};

constexpr int FILTER_MAX_TERMS = 3;

struct FilterPlan {
  uint32_t kind;
  uint32_t nterms;                  // FK_LUT: 1..3
  uint32_t t_off[FILTER_MAX_TERMS]; // FK_LUT: 0..8
  uint32_t a_off[2];                // FK_ANCHOR2: offsets (0, 1..12) of the two compared bytes
  uint32_t a_chr[2];                // ... and the bytes, splatted over a word
  uint32_t h4_terms;                // 0..3: hashed-predictor terms (min_ >= 4): position k survives iff for t < h4_terms
                                    //   bit (3 + t) of pmh[g(k + h4_shift + 3 + t)] is clear, g = the 12-bit rolling hash of
                                    //   the 4 bytes ending at its argument (Pattern::predict_match steps 3, 4, 5)
  uint32_t h4_shift;                // the predictor starts at k + h4_shift (1 for the CHAR_PMH routine, else 0)
  uint32_t pm2;                     // 1: PM4 two-byte term (min_ < 4): with q7 q6 = bits 7, 6 of pma[c0] and q5 q4 = bits 5, 4
                                    //   of pma[hash(c0, c1)], position k fails iff q7 & q5 & (q4 | q6) — the part of
                                    //   Pattern::predict_match's PM4 formula that the first two bytes decide
  uint32_t pm2_shift;               // the predictor starts at k + pm2_shift (1 for the CHAR_PMA routine, else 0)
  uint32_t est_pass_ppm;            // planner's estimate of the survivor rate (informational)
  uint32_t lut[256];                // FK_LUT: bit 8*t of lut[c] set = byte c FAILS term t (the planes are byte-
                                    // aligned so that eight positions accumulate in one register: acc = 2*acc + lut[c])
};

} // namespace ugx
