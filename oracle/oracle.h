/*
 * oracle/oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference's buffer-scan path
 * (reflex::Matcher::match(FIND) + the advance_* prefilters + the three caller
 * loops of Grep::search) over the same compiled-pattern tables the product
 * consumes (include/ugrep_b200.h).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this; the product never does.
 *
 * Parity pin: tests/test_oracle_vs_reference.py checks this restatement against
 * the reference's own golden files (tests/out/ *.out) and against the unmodified
 * reference built into oracle/_ref (ugrep CLI and libreflex in-place scans).
 */
#ifndef UGX_ORACLE_H
#define UGX_ORACLE_H

#include "../include/ugrep_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_pattern ora_pattern;

int  ora_pattern_create(const uint32_t *opc, uint32_t nop, const ugx_prefilter *pf,
                        uint32_t matcher_flags, ora_pattern **out);
int  ora_pattern_load(const char *path, ora_pattern **out);
void ora_pattern_destroy(ora_pattern *p);
int  ora_advance_kind(const ora_pattern *p); /* UGX_ADV_* */

/* the three caller loops (src/ugrep.cpp:10567-10586, :10536-10566, :10857-11047) */
int  ora_count_lines(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint64_t *count);
int  ora_count_matches(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint64_t *count);
int  ora_find_all(const ora_pattern *p, const uint8_t *buf, uint64_t n,
                  uint64_t base_offset, uint64_t base_line,
                  ugx_match *out, uint64_t cap, uint64_t *n_out);
uint64_t ora_count_newlines(const uint8_t *buf, uint64_t n);
/* reflex::isutf8 (lib/simd.cpp:169-421): ugrep's binary-file test is !isutf8 */
int  ora_isutf8(const uint8_t *buf, uint64_t n);
int  ora_has_nul(const uint8_t *buf, uint64_t n);

/* prefilter alone: bit k of bitmap (little-endian within bytes) = position k is a candidate */
int  ora_candidates(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint8_t *bitmap);
/* one anchored DFA attempt at position k: returns accept index (0 = none), *len = match length */
int  ora_match_at(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint64_t k, uint64_t *len);

#ifdef __cplusplus
}
#endif
#endif
