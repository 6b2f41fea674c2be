/*
 * ugrep_b200.h — C ABI of the B200-native buffer-scan path for ugrep.
 *
 * This is the drop-in boundary.  The reference has no C plugin
 * ABI; its boundary is the C++ pair reflex::Pattern / reflex::Matcher.  The
 * functions below are what a patched reflex::Matcher binds (see
 * INTEGRATION.md): the compiled pattern tables go in once, a buffer goes in per
 * file, and counts / (line, offset, length, accept) records come back.
 *
 * Everything is plain pointers and sizes.  There is NO CPU fallback: every scan
 * entry point runs sm_100a CUDA kernels or returns a non-zero status.
 *
 * Reference interfaces replaced (all paths into /root/reference):
 *   ugx_pattern_create    reflex::Pattern tables as consumed by reflex::Matcher
 *                         (include/reflex/pattern.h:1288-1334; the (code, pred)
 *                         constructor pattern.h:151-159, loader lib/pattern.cpp:199-274)
 *   ugx_count_lines       `ugrep -c`    loop, src/ugrep.cpp:10567-10586  (find + skip('\n'))
 *   ugx_count_matches     `ugrep -c -o` loop, src/ugrep.cpp:10536-10566  (find)
 *   ugx_find_all          `ugrep -o [-n -b]` loop, src/ugrep.cpp:10857-11047 and the
 *                         state Output::header reads (lineno(), first(), begin(), size();
 *                         include/reflex/absmatcher.h:695-766, :901-905)
 *   ugx_count_newlines    reflex::nlcount, lib/simd.cpp:62-166
 * each of which is a loop over reflex::Matcher::match(FIND), lib/matcher.cpp:42-750,
 * with the prefilters of lib/matcher.cpp:797-3549 / lib/matcher_avx2.cpp.
 */
#ifndef UGREP_B200_H
#define UGREP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGX_ABI_VERSION 2

/* status codes (0 = ok; nothing ever falls back to the CPU) */
enum {
  UGX_OK            = 0,
  UGX_E_INVALID     = 1, /* bad argument / malformed pattern tables */
  UGX_E_UNSUPPORTED = 2, /* pattern uses a feature outside the path's scope (HEAD/TAIL lookahead, REDO, '\n' transitions, ...) */
  UGX_E_CUDA        = 3, /* CUDA runtime error, see ugx_last_error() */
  UGX_E_NOMEM       = 4, /* host or device allocation failed */
  UGX_E_OVERFLOW    = 5, /* caller's record buffer too small: *n_out holds the required count */
  UGX_E_IO          = 6  /* pattern file unreadable */
};

/* reflex::Matcher option letters (include/reflex/absmatcher.h:354-388) */
#define UGX_OPT_N 0x01u /* "N": accept empty matches (ugrep -Y) */
#define UGX_OPT_W 0x02u /* "W": whole-word matching (ugrep -w) */

#define UGX_BTAP 2048 /* reflex::Pattern::Const::BTAP */
#define UGX_HASH 4096 /* reflex::Pattern::Const::HASH */

/*
 * The prefilter block of a compiled reflex::Pattern: a field-for-field image of
 * the members reflex::Matcher reads (include/reflex/pattern.h:1305-1334).
 * bit/tap/pma/pmh hold the in-memory values (a 0 bit = "may match").
 */
typedef struct ugx_prefilter {
  uint32_t len;  /* len_: length of the literal prefix in chr[] (0 = none) */
  uint32_t min;  /* min_: minimum length of what follows the prefix, 0..8 */
  uint32_t pin;  /* pin_: number of needle bytes per needle position, 0..16 */
  uint32_t lcp;  /* lcp_: primary needle position */
  uint32_t lcs;  /* lcs_: secondary needle position */
  uint32_t bmd;  /* bmd_: Boyer-Moore distance (>0 selects the B-M routines) */
  uint32_t npy;  /* npy_: bitap entropy */
  uint32_t one;  /* one_: pattern is the single literal chr[0..len) */
  uint32_t bol;  /* bol_: every alternative is anchored with ^ */
  uint32_t lbk;  /* lbk_: look-back distance, 0xffff = unbounded, 0 = none */
  uint32_t lbm;  /* lbm_: minimum look-back distance */
  uint32_t cut;  /* cut_: DFA s-t cut depth (informational) */
  uint8_t  chr[256];      /* chr_: literal prefix, or 2*pin needle bytes */
  uint8_t  bit[256];      /* bit_ */
  uint8_t  tap[UGX_BTAP]; /* tap_: bitap hashed byte pairs */
  uint8_t  pma[UGX_HASH]; /* pma_: PM4 predictor (min < 4) */
  uint8_t  pmh[UGX_HASH]; /* pmh_: hashed Bloom predictor (min >= 4) */
  uint8_t  cbk[32];       /* cbk_ as a 256-bit little-endian bitset */
  uint8_t  fst[32];       /* fst_ as a 256-bit little-endian bitset */
  uint8_t  bms[256];      /* bms_: Boyer-Moore shifts */
} ugx_prefilter;

/* UGXP container: header, ugx_prefilter, nop opcode words, regex text (informational) */
#define UGX_FILE_MAGIC "UGXP\1\0\0\0"
typedef struct ugx_file_header {
  char     magic[8];
  uint32_t nop;            /* number of 32-bit opcode words */
  uint32_t regex_len;      /* bytes of regex text after the opcode words */
  uint32_t prefilter_size; /* sizeof(ugx_prefilter) at write time */
  uint32_t matcher_flags;  /* UGX_OPT_* the pattern was meant to be used with */
} ugx_file_header;

/* one match, in input order: what lineno()/first()/size()/accept() return for it */
typedef struct ugx_match {
  uint64_t line;   /* 1-based line number of the match start (+ base_line) */
  uint64_t offset; /* 0-based byte offset of the match start (+ base_offset) */
  uint32_t len;    /* length in bytes */
  uint32_t cap;    /* accept index: 1-based alternative number */
} ugx_match;

/* facts about an uploaded pattern */
typedef struct ugx_pattern_info {
  uint32_t nop;          /* opcode words */
  uint32_t states;       /* reachable DFA states after flattening */
  uint32_t classes;      /* byte equivalence classes */
  uint32_t table_bytes;  /* dense transition table size */
  uint32_t table_in_smem;/* 1 if the scan kernels stage the table in shared memory */
  uint32_t advance;      /* UGX_ADV_* prefilter routine selected (init_advance, lib/matcher.cpp:797-954) */
  uint32_t has_meta;     /* DFA has META (anchor / word boundary) edges */
  uint32_t lookback;     /* lbk != 0 */
} ugx_pattern_info;

/* prefilter routine families of Matcher::init_advance (lib/matcher.cpp:797-954) */
enum {
  UGX_ADV_NONE = 0,
  UGX_ADV_PIN1_ONE, UGX_ADV_PIN1_PMA, UGX_ADV_PIN1_PMH,
  UGX_ADV_PIN_ONE,  UGX_ADV_PIN_PMA,  UGX_ADV_PIN_PMH,
  UGX_ADV_MIN1, UGX_ADV_MIN2, UGX_ADV_MIN3, UGX_ADV_MIN4, UGX_ADV_PMA,
  UGX_ADV_CHAR, UGX_ADV_CHAR_PMA, UGX_ADV_CHAR_PMH,
  UGX_ADV_STRING, UGX_ADV_STRING_PMA, UGX_ADV_STRING_PMH
};

/* what the host-side export derives from a compiled pattern, without touching a device: the dense DFA's shape and the
 * first-stage filter the position-parallel kernels use (csrc/filter_plan.hpp).  For tests and diagnostics. */
typedef struct ugx_plan_info {
  uint32_t states, classes, table_bytes;
  uint32_t first_acc, first_leaf; /* state numbering: [1, first_acc) plain, [first_acc, first_leaf) accepting, rest leaves */
  uint32_t max_match_len;         /* longest match in bytes, 0xffffffff = unbounded (cyclic DFA) */
  uint32_t advance;               /* UGX_ADV_* */
  uint32_t has_meta, newline_live;
  uint32_t kind;                  /* 0 all survive, 2 two literal anchors, 3 byte-set terms */
  uint32_t nterms, t_off[3];      /* byte-set terms: byte (k + t_off[t]) must not have bit 8*t set in lut[] */
  uint32_t a_off[2], a_chr[2];    /* literal anchors: offsets and bytes */
  uint32_t h4_terms, h4_shift;    /* hashed-predictor terms (steps 3.. of predict_match PMH) */
  uint32_t pm2, pm2_shift;        /* PM4 two-byte term */
  uint32_t lut[256];
  uint32_t covers;                /* proven by enumeration: every position that starts a non-empty match passes the
                                     routine's candidate test (so `-c` may skip the test); 0 = not proven or not so */
} ugx_plan_info;

enum { UGX_MODE_LINES = 0, UGX_MODE_MATCHES = 1, UGX_MODE_RECORDS = 2 }; /* ugrep -c | -c -o | -o -n -b */

typedef struct ugx_pattern ugx_pattern; /* immutable once created; shareable between host threads */
typedef struct ugx_scanner ugx_scanner; /* per host thread / per stream scratch (counters, record arena) */

/* totals every scan reports */
typedef struct ugx_totals {
  uint64_t matches;        /* count-lines: matching lines; otherwise: matches */
  uint64_t newlines;       /* '\n' bytes in the buffer (line-number base for the next shard); valid when flags has
                              UGX_TOT_NEWLINES: every call except ugx_count_lines on its streaming kernels, which
                              count newlines only with the scanner option "count_newlines" */
  uint64_t flags;          /* UGX_TOT_* */
  float    kernel_ms;      /* device time of the scan kernels (CUDA events on the scan stream) */
  uint32_t launches;       /* kernels launched by this call */
  uint32_t kernel;         /* UGX_K_*: the scan kernel that did the work (ugx_kernel_name) */
} ugx_totals;

#define UGX_TOT_NEWLINES      1u /* `newlines` was counted by this call */
#define UGX_TOT_SPAN_HANDOVER 2u /* the span kernels could not vouch for this buffer (a match longer than a window across
                                    a region start in a line without newlines, a look-back run or match beyond their
                                    bounds, a failed attempt at the very end): the line-at-a-time kernels did the scan */

/* scan kernels (reported in ugx_totals.kernel) */
enum { UGX_K_NONE = 0, UGX_K_STREAM_LITERAL = 1, UGX_K_STREAM_DFA = 2, UGX_K_TILE_ANY = 3, UGX_K_LINE_SCAN = 4,
       UGX_K_RECORDS = 5, UGX_K_NEWLINES = 6, UGX_K_MATCH_LINES = 7, UGX_K_SPAN = 8, UGX_K_BATCH = 9 };

const char *ugx_last_error(void);
const char *ugx_kernel_name(uint32_t id);
int  ugx_abi_version(void);

/* pattern: flatten the opcode table to a dense class-compressed DFA and upload it with the prefilter tables */
int  ugx_pattern_create(const uint32_t *opc, uint32_t nop, const ugx_prefilter *pf,
                        uint32_t matcher_flags, int device, ugx_pattern **out);
int  ugx_pattern_load(const char *path, int device, ugx_pattern **out);
/* host only: the compiled form of ONE fixed string (`ugrep -F 'literal'`): the opcode words and prefilter fields the
 * reference's pattern compiler produces for it (lib/pattern.cpp:2823-3063, 4286-4340, 510-598), byte for byte, so that
 * a literal search needs no reference binary.  *nop receives the word count (UGX_E_OVERFLOW when cap is too small);
 * UGX_E_UNSUPPORTED for an empty literal, 255 bytes or more, or one that holds NUL / CR / LF.  The general regex and
 * word-list compiler is not part of this library. */
int  ugx_compile_literal(const uint8_t *literal, uint32_t len, uint32_t *opc, uint32_t cap, uint32_t *nop,
                         ugx_prefilter *pf);
/* host only: the same for a LIST of fixed strings (`ugrep -F -f words.txt`, `-F -e A -e B`; accept index = 1-based
 * position in the list): the tree DFA's opcode words and the needle / bitap / hashed-predictor tables, byte for byte
 * what the reference compiles (lib/pattern.cpp:798-866, 3812-4639, 331-598, 2764-3063).  UGX_E_UNSUPPORTED: an empty
 * string, NUL / CR / LF in a string, or a list for which the reference's DFA analysis makes a cut (look-back search:
 * that part of the analysis is not restated). */
int  ugx_compile_words(const uint8_t *const *words, const uint32_t *lens, uint32_t nwords, uint32_t *opc, uint32_t cap,
                       uint32_t *nop, ugx_prefilter *pf);
/* the same with options: UGX_COMPILE_ICASE = `ugrep -F -i -f words.txt` (the strings are lowered as they enter the tree
 * and every edge on a lowercase ASCII letter gets an uppercase twin, lib/pattern.cpp:286-311, 834) */
enum { UGX_COMPILE_ICASE = 1 };
int  ugx_compile_words_ex(const uint8_t *const *words, const uint32_t *lens, uint32_t nwords, uint32_t options,
                          uint32_t *opc, uint32_t cap, uint32_t *nop, ugx_prefilter *pf);
/* host only: a REGEX that is an alternation of plain strings — `ugrep [-i] -e 'foo\.bar|baz'` without -F.  The
 * reference's parser puts such a regex into the same tree DFA as a -F list (lib/pattern.cpp:286-311, 798-866), so this
 * is ugx_compile_words_ex on the alternatives.  Accepted: bytes other than the operators \ . [ ] ( ) { } * + ? | ^ $,
 * top-level `|`, a backslash before an operator or one of ! " # % & ' , - / : ; @ ` (that character), \t \f \v \a,
 * \Q...\E, a leading (?i) (= UGX_COMPILE_ICASE); with UGX_COMPILE_ICASE no byte >= 0x80 outside \Q...\E.  Anything else (classes, groups, repeats, anchors,
 * \d \w \b \xHH ...) returns UGX_E_UNSUPPORTED: the regex compiler proper is not part of this library.  Alternative i (1-based) is accept index i. */
int  ugx_compile_plain_regex(const uint8_t *regex, uint32_t len, uint32_t options, uint32_t *opc, uint32_t cap,
                             uint32_t *nop, ugx_prefilter *pf);
/* host only (no device needed): DFA export + filter plan of a compiled pattern */
int  ugx_plan_describe(const uint32_t *opc, uint32_t nop, const ugx_prefilter *pf, uint32_t matcher_flags,
                       ugx_plan_info *out);
/* host only: the k-gram viability tables of the scan kernels (csrc/pattern_host.hpp).  *k = bytes looked at (0 = no
 * table), *stride = bits per state code, t01[256] / t23[256] = the pre-multiplied byte ids, pair[cap_pair] = the
 * two-byte state codes, bits[cap_words] = the bit table; *npair / *words = their sizes */
int  ugx_viability_describe(const uint32_t *opc, uint32_t nop, uint32_t *k, uint32_t *stride, uint32_t *t01, uint32_t *t23,
                            uint8_t *pair, uint32_t cap_pair, uint32_t *npair,
                            uint32_t *bits, uint32_t cap_words, uint32_t *words);
int  ugx_pattern_info_get(const ugx_pattern *p, ugx_pattern_info *info);
void ugx_pattern_destroy(ugx_pattern *p);

/* scanner: owns a stream-ordered scratch arena on `device`; `stream` is a cudaStream_t (NULL = default stream) */
int  ugx_scanner_create(int device, void *stream, ugx_scanner **out);
void ugx_scanner_destroy(ugx_scanner *s);
/* options (tests and A/B timing; defaults 0):
 *   "force_generic"   every scan takes the generic line-scan kernel
 *   "legacy_any"      `-c` takes the tile-synchronous kernel instead of the streaming one
 *   "stream_dfa"      `-c` of every eligible DFA pattern takes the streaming kernel
 *   "count_newlines"  the streaming `-c` kernels also count newlines (totals.newlines)
 *   "match_lines"     counting takes the position-parallel-attempt kernel (match_lines.cu) instead of the line scan
 *   "two_pass_records" records by a count pass + an emit pass instead of the single-pass staging form
 *   "no_pipeline"     host buffers: one copy, then the scan (default: chunked copy overlapped with the scan)
 *   "no_feeder"       pageable host buffers take one plain cudaMemcpyAsync instead of the feeder threads
 *   "no_span"         counting matches / records take the line-at-a-time kernels instead of the span kernels */
int  ugx_scanner_set_option(ugx_scanner *s, const char *name, int value);

/*
 * Scans.  `buf` is nbytes of text; it may be a device pointer (scanned in place)
 * or a host pointer (staged to the device inside the call, chunked and
 * overlapped with the scan).  The buffer is one file or one line-aligned shard.
 */
int  ugx_count_lines(ugx_scanner *s, const ugx_pattern *p, const void *buf, uint64_t nbytes,
                     ugx_totals *totals);
int  ugx_count_matches(ugx_scanner *s, const ugx_pattern *p, const void *buf, uint64_t nbytes,
                       ugx_totals *totals);
/* records are written to the host array `out` (capacity `cap`), in input order */
int  ugx_find_all(ugx_scanner *s, const ugx_pattern *p, const void *buf, uint64_t nbytes,
                  uint64_t base_offset, uint64_t base_line,
                  ugx_match *out, uint64_t cap, uint64_t *n_out, ugx_totals *totals);
/* records stay on the device: *dev_out is owned by the scanner and valid until its next call */
int  ugx_find_all_device(ugx_scanner *s, const ugx_pattern *p, const void *buf, uint64_t nbytes,
                         uint64_t base_offset, uint64_t base_line,
                         const ugx_match **dev_out, uint64_t *n_out, ugx_totals *totals);
/* copy `count` records of the last ugx_find_all_device result, starting at record `first`, to the host */
int  ugx_scanner_fetch(ugx_scanner *s, ugx_match *out, uint64_t first, uint64_t count);
int  ugx_count_newlines(ugx_scanner *s, const void *buf, uint64_t nbytes, ugx_totals *totals);

/*
 * Many files in one launch: what the reference's per-file job queue feeds its workers (GrepMaster::submit /
 * GrepWorker::execute, src/ugrep.cpp:4295-4432).  The files lie in one buffer (host or device), file i at
 * buf[begins[i], begins[i] + lens[i]) with every begins[i] a multiple of 16 (the bytes between files are never read as
 * text); counts[i] receives what `ugrep -c` (UGX_MODE_LINES) or `ugrep -c -o` (UGX_MODE_MATCHES) prints for file i
 * scanned on its own: each file keeps its own end of buffer.  totals->matches is the sum.
 */
int  ugx_count_batch(ugx_scanner *s, const ugx_pattern *p, const void *buf, uint64_t nbytes,
                     const uint64_t *begins, const uint64_t *lens, uint64_t nfiles, int mode,
                     uint64_t *counts, ugx_totals *totals);

/* what ugrep asks about a file before it prints from it (Grep::init_is_binary, src/ugrep.cpp:3998-4017):
 * is_utf8 = reflex::isutf8 (lib/simd.cpp:169-421: valid UTF-8 structure, no NUL), has_nul = memchr(buf, 0, n) != NULL;
 * is_binary() is !is_utf8 by default and has_nul with -U (src/ugrep.cpp:699-711) */
typedef struct ugx_text_info {
  uint32_t is_utf8;
  uint32_t has_nul;
  float    kernel_ms;
  uint32_t launches;
} ugx_text_info;
int  ugx_check_text(ugx_scanner *s, const void *buf, uint64_t nbytes, ugx_text_info *out);

/*
 * One process, several GPUs: the buffer is cut into one line-aligned shard per device (forward from
 * n*r/N to the next newline), every device scans its shard, and the per-shard {matches, newlines} give the totals and
 * the record / line-number bases.  Replaces the reference's per-file worker pool (GrepMaster / GrepWorker,
 * src/ugrep.cpp:4118-4432) for one large input.  `devices` may name a device more than once.
 */
typedef struct ugx_sharded ugx_sharded;
typedef struct ugx_shard {
  int32_t  device;
  uint32_t reserved;
  uint64_t begin, end;       /* the shard is buf[begin, end) */
  uint64_t matches, newlines;
  uint64_t line_base;        /* newlines before the shard */
  uint64_t record_base;      /* matches before the shard */
  float    kernel_ms;
  uint32_t reserved2;
} ugx_shard;
int  ugx_sharded_create(const uint32_t *opc, uint32_t nop, const ugx_prefilter *pf, uint32_t matcher_flags,
                        const int *devices, int ndev, ugx_sharded **out);
void ugx_sharded_destroy(ugx_sharded *s);
/* "pin": page-lock the caller's buffer during a scan (pays when the same buffer is scanned repeatedly);
 * every other name is passed to each device's scanner (ugx_scanner_set_option) */
int  ugx_sharded_set_option(ugx_sharded *s, const char *name, int value);
/* host_buf: host memory.  mode UGX_MODE_RECORDS writes the records of all shards to out[cap] in input order, offsets
 * and line numbers global; shards[ndev] (optional) receives what every device did */
int  ugx_sharded_scan(ugx_sharded *s, const void *host_buf, uint64_t nbytes, int mode, ugx_match *out, uint64_t cap,
                      uint64_t *n_out, ugx_totals *totals, ugx_shard *shards);
const char *ugx_sharded_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* UGREP_B200_H */
