/*
 * oracle/oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference's buffer-scan path, written from the
 * reference's behaviour, not from its text.  Each function names the reference
 * code it follows (paths into /root/reference).
 *
 * Scope of the restatement (same as the product): in-place one-pass buffers
 * (AbstractMatcher::buffer, include/reflex/absmatcher.h:542-591), method FIND,
 * patterns without HEAD/TAIL/REDO opcodes, matcher options W and N honoured
 * (N = ugrep -Y, implied by ^... and ...$ patterns, src/cnf.hpp:199-203: find()
 * then returns empty matches, lib/matcher.cpp:681-728, and a pattern whose
 * minimum length is 0 is searched without a prefilter, :804).
 *
 * The prefilters are restated in POSITION-LOCAL form: cand(k) says whether the
 * reference's advance routine can stop at byte k.  Where the reference's SIMD
 * main loop and its scalar tail apply different tests to the same position
 * (which of the two runs depends on the alignment of the call, and on the
 * 256 KiB stream window in the CLI), cand(k) is the main-loop test wherever
 * all bytes it reads exist, and the routine's own end-of-buffer rule in the
 * last few bytes.  DESIGN.md ("candidate predicates") lists them one by one.
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OP_HALT 0x00FFFFFFu
#define IDX_HALT 0xFFFFu
#define IDX_LONG 0xFFFEu
#define NONE 0xFFFFFFFFu
#define CH_EOF (-1)

struct ora_pattern {
  uint32_t *opc;
  uint32_t nop;
  ugx_prefilter pf;
  uint32_t flags;
  int adv;
  uint8_t pin_a[256]; /* needle set at position lcp */
  uint8_t pin_b[256]; /* needle set at position lcs */
};

static const int word_ranges[] = {
#include "../ugrep_b200/csrc/word_ranges.inc"
};
#define N_WORD_RANGES ((int)(sizeof(word_ranges) / sizeof(int) / 2))

/* reflex::Matcher::iswword, include/reflex/matcher.h:457-1192 (binary search over [lo,hi] ranges) */
static int is_word_cp(int c)
{
  int lo = 0, hi = N_WORD_RANGES - 1;
  while (lo <= hi)
  {
    int mid = (lo + hi) / 2;
    if (c < word_ranges[2 * mid])
      hi = mid - 1;
    else if (c > word_ranges[2 * mid + 1])
      lo = mid + 1;
    else
      return 1;
  }
  return 0;
}

static int is_alnum_ascii(int c)
{
  return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
}

/* ---- scan state: the subset of AbstractMatcher fields the path touches (absmatcher.h:1632-1660) ---- */
typedef struct {
  const ora_pattern *p;
  const uint8_t *b;
  size_t end; /* end_ : number of text bytes */
  size_t cur; /* cur_ */
  size_t pos; /* pos_ */
  size_t txt; /* txt_ - buf_ */
  size_t len; /* len_ */
  int got;    /* got_ : byte before cur_, '\n' at offset 0 (set_current, absmatcher.h:1571-1580) */
  uint32_t cap;
} scan_t;

static inline int byte_at(const scan_t *m, size_t i)
{
  return i < m->end ? m->b[i] : 0; /* the slot after the text is the caller's NUL (src/ugrep.cpp:3939) */
}

/* reflex::utf8(), include/reflex/utf8.h:138-215 (restricted UTF-8, invalid -> U+FFFD) */
static int decode_utf8(const scan_t *m, size_t i)
{
  int c = byte_at(m, i);
  if (c < 0x80)
    return c;
  int c1 = byte_at(m, i + 1);
  if (c < 0xC0 || (c == 0xC0 && c1 != 0x80) || c == 0xC1 || (c1 & 0xC0) != 0x80)
    return 0xFFFD;
  c1 &= 0x3F;
  if (c < 0xE0)
    return ((c & 0x1F) << 6) | c1;
  int c2 = byte_at(m, i + 2);
  if ((c == 0xE0 && c1 < 0x20) || (c2 & 0xC0) != 0x80)
    return 0xFFFD;
  c2 &= 0x3F;
  if (c < 0xF0)
    return ((c & 0x0F) << 12) | (c1 << 6) | c2;
  int c3 = byte_at(m, i + 3);
  if ((c == 0xF0 && c1 < 0x10) || (c == 0xF4 && c1 >= 0x10) || c >= 0xF5 || (c3 & 0xC0) != 0x80)
    return 0xFFFD;
  return ((c & 0x07) << 18) | (c1 << 12) | (c2 << 6) | (c3 & 0x3F);
}

static inline void set_current(scan_t *m, size_t loc)
{
  m->pos = m->cur = loc;
  m->got = loc > 0 ? m->b[loc - 1] : '\n';
}

static inline int at_end(const scan_t *m) { return m->pos >= m->end; }
static inline int get_ch(scan_t *m) { return m->pos < m->end ? m->b[m->pos++] : CH_EOF; }
static inline int peek_ch(const scan_t *m) { return m->pos < m->end ? m->b[m->pos] : CH_EOF; }

/* ---- word boundary predicates, include/reflex/matcher.h:1194-1319 ---- */
static int at_wb(const scan_t *m)
{
  int c = m->got;
  if (c == '\n')
    return 1;
  if (c == '_')
    return 0;
  if ((c & 0xC0) == 0x80 && m->cur > 0)
  {
    size_t k = m->cur - 1;
    if (k > 0 && (m->b[--k] & 0xC0) == 0x80)
      if (k > 0 && (m->b[--k] & 0xC0) == 0x80)
        if (k > 0)
          --k;
    return !is_word_cp(decode_utf8(m, k));
  }
  return !is_alnum_ascii(c);
}

static int at_we(const scan_t *m, int c, size_t k)
{
  if (c == CH_EOF)
    return 1;
  if (c == '_')
    return 0;
  if ((c & 0xC0) == 0xC0)
    return !is_word_cp(decode_utf8(m, k));
  return !is_alnum_ascii(c);
}

static int at_bw(const scan_t *m)
{
  size_t i = m->txt + m->len;
  int c = byte_at(m, i);
  if (c == '_')
    return 1;
  if ((c & 0xC0) == 0xC0)
    return is_word_cp(decode_utf8(m, i));
  return is_alnum_ascii(c);
}

static int at_ew(const scan_t *m, int c)
{
  size_t k = m->pos + (c == CH_EOF);
  c = k > 1 ? m->b[k - 2] : m->got;
  if (c == '\n')
    return 0;
  if (c == '_')
    return 1;
  if ((c & 0xC0) == 0x80 && k > 2)
  {
    k -= 3;
    if ((m->b[k] & 0xC0) == 0x80)
      if (k > 0 && (m->b[--k] & 0xC0) == 0x80)
        if (k > 0)
          --k;
    return is_word_cp(decode_utf8(m, k));
  }
  return is_alnum_ascii(c);
}

/* ---- predictors, include/reflex/pattern.h:366-401, hashes :1274-1282 ---- */
static inline uint32_t hash3(uint32_t h, uint32_t b) { return ((h << 3) ^ b) & (UGX_HASH - 1); }
static inline uint32_t bihash(uint32_t a, uint32_t b) { return (a ^ (b << 6)) & (UGX_BTAP - 1); }

static int pm4(const scan_t *m, size_t k)
{
  const uint8_t *pma = m->p->pf.pma;
  uint32_t c0 = byte_at(m, k), c1 = byte_at(m, k + 1), c2 = byte_at(m, k + 2), c3 = byte_at(m, k + 3);
  uint32_t h1 = hash3(c0, c1), h2 = hash3(h1, c2), h3 = hash3(h2, c3);
  uint8_t q = (pma[c0] & 0xc0) | (pma[h1] & 0x30) | (pma[h2] & 0x0c) | (pma[h3] & 0x03);
  uint8_t r = (uint8_t)(((((((q >> 2) | q) >> 2) | q) >> 1) | q));
  return r != 0xff;
}

static int pmh(const scan_t *m, size_t k, size_t n)
{
  const uint8_t *t = m->p->pf.pmh;
  uint32_t h = byte_at(m, k);
  uint32_t f = t[h] & 1;
  uint32_t bit = 2;
  for (size_t j = 1; j < n; ++j, bit <<= 1)
  {
    h = hash3(h, byte_at(m, k + j));
    f |= t[h] & bit;
    if (j == 3 && f != 0)
      return 0;
  }
  return f == 0;
}

static inline int tapbit(const scan_t *m, size_t k, unsigned j)
{
  return (m->p->pf.tap[bihash(byte_at(m, k), byte_at(m, k + 1))] >> j) & 1;
}

static int literal_at(const scan_t *m, size_t k)
{
  const ugx_prefilter *pf = &m->p->pf;
  if (k + pf->len > m->end)
    return 0;
  return memcmp(m->b + k, pf->chr, pf->len) == 0;
}

/* Matcher::init_advance, lib/matcher.cpp:797-954 (+ the AVX2/AVX512BW overrides, same families) */
static int select_advance(const ugx_prefilter *pf, uint32_t flags)
{
  if (pf->len == 0)
  {
    if (pf->min == 0 && (flags & UGX_OPT_N))
      return UGX_ADV_NONE;
    if (pf->pin == 1)
      return pf->min < 2 ? UGX_ADV_PIN1_ONE : pf->min < 4 ? UGX_ADV_PIN1_PMA : UGX_ADV_PIN1_PMH;
    if ((pf->pin >= 2 && pf->pin <= 8) || pf->pin == 16)
      return pf->min < 2 ? UGX_ADV_PIN_ONE : pf->min < 4 ? UGX_ADV_PIN_PMA : UGX_ADV_PIN_PMH;
    switch (pf->min)
    {
      case 0:
      case 1: return pf->npy <= 33 ? UGX_ADV_MIN1 : UGX_ADV_PMA;
      case 2: return pf->npy <= 36 ? UGX_ADV_MIN2 : UGX_ADV_PMA;
      case 3: return pf->npy <= 47 ? UGX_ADV_MIN3 : UGX_ADV_PMA;
      default: return UGX_ADV_MIN4;
    }
  }
  if (pf->len == 1)
    return pf->min == 0 ? UGX_ADV_CHAR : pf->min < 4 ? UGX_ADV_CHAR_PMA : UGX_ADV_CHAR_PMH;
  /* chars<2>, chars<3>, string and string_bm share one predicate: the literal, then the predictor */
  return pf->min == 0 ? UGX_ADV_STRING : pf->min < 4 ? UGX_ADV_STRING_PMA : UGX_ADV_STRING_PMH;
}

/*
 * cand(k): can the reference's advance routine stop at k?  (0 <= k < end)
 * One case per routine family; the line numbers are the routine restated.
 */
static int cand(const scan_t *m, size_t k)
{
  const ugx_prefilter *pf = &m->p->pf;
  const size_t end = m->end;
  const size_t min = pf->min, len = pf->len, lcp = pf->lcp, lcs = pf->lcs;
  switch (m->p->adv)
  {
    case UGX_ADV_PIN1_ONE: /* lib/matcher.cpp:963-990 */
      return m->b[k] == pf->chr[0] && (k + 4 > end || pm4(m, k));
    case UGX_ADV_PIN1_PMA: /* lib/matcher.cpp:993-1109 */
      if (k + lcp >= end || m->b[k + lcp] != pf->chr[0])
        return 0;
      return k + 4 > end || (byte_at(m, k + lcs) == pf->chr[1] && pm4(m, k));
    case UGX_ADV_PIN1_PMH: /* lib/matcher.cpp:1112-1228 */
      if (k + lcp >= end || m->b[k + lcp] != pf->chr[0])
        return 0;
      return k + min > end || (byte_at(m, k + lcs) == pf->chr[1] && pmh(m, k, min));
    case UGX_ADV_PIN_ONE: /* lib/matcher.cpp:1233-1276 */
      if (k + 4 > end)
        return 1;
      return m->p->pin_a[m->b[k]] && pm4(m, k);
    case UGX_ADV_PIN_PMA: /* lib/matcher.cpp:1370-1422 */
      if (k + min > end)
        return 0;
      if (k + 4 > end)
        return 1;
      return m->p->pin_a[m->b[k + lcp]] && m->p->pin_b[m->b[k + lcs]] && pm4(m, k);
    case UGX_ADV_PIN_PMH: /* lib/matcher.cpp:1424-1473 */
      if (k + min > end)
        return 0;
      return m->p->pin_a[m->b[k + lcp]] && m->p->pin_b[m->b[k + lcs]] && pmh(m, k, min);
    case UGX_ADV_MIN1: /* lib/matcher.cpp:2248-2304 */
      if (tapbit(m, k, 0))
        return 0;
      return k + 4 >= end || pm4(m, k);
    case UGX_ADV_MIN2: /* lib/matcher.cpp:2307-2345 */
      if (k + 2 > end)
        return 0;
      if (tapbit(m, k, 0) || tapbit(m, k + 1, 1))
        return 0;
      return k + 5 > end || pm4(m, k);
    case UGX_ADV_MIN3: /* lib/matcher.cpp:2348-2386 */
      if (k + 3 > end)
        return 0;
      if (tapbit(m, k, 0) || tapbit(m, k + 1, 1) || tapbit(m, k + 2, 2))
        return 0;
      return k + 5 > end || pm4(m, k);
    case UGX_ADV_MIN4: /* lib/matcher.cpp:2389-2462 */
      if (k + min > end)
        return 0;
      for (unsigned j = 0; j < min; ++j)
        if (tapbit(m, k + j, j))
          return 0;
      return pmh(m, k, min);
    case UGX_ADV_PMA: /* lib/matcher.cpp:2465-2489, ending in advance_pattern_min1 */
      if (k + 7 <= end)
        return pm4(m, k);
      if (tapbit(m, k, 0))
        return 0;
      return k + 4 >= end || pm4(m, k);
    case UGX_ADV_CHAR: /* lib/matcher.cpp:2492-2512 */
      return m->b[k] == pf->chr[0];
    case UGX_ADV_CHAR_PMA: /* lib/matcher.cpp:2515-2542 */
      return m->b[k] == pf->chr[0] && (k + 5 > end || pm4(m, k + 1));
    case UGX_ADV_CHAR_PMH: /* lib/matcher.cpp:2545-2573 */
      return m->b[k] == pf->chr[0] && (k + 1 + min > end || pmh(m, k + 1, min));
    case UGX_ADV_STRING: /* lib/matcher.cpp:2577-2693, :2962-3025, :3377-3430 */
      return literal_at(m, k);
    case UGX_ADV_STRING_PMA: /* lib/matcher.cpp:2697-2826, :3028-3098, :3433-3489 */
      if (k + len + min > end || !literal_at(m, k))
        return 0;
      return k + len + 4 > end || pm4(m, k + len);
    case UGX_ADV_STRING_PMH: /* lib/matcher.cpp:2830-2959, :3101-3171, :3492-3549 */
      if (k + len + min > end || !literal_at(m, k))
        return 0;
      return pmh(m, k + len, min);
    default:
      return 0;
  }
}

/* (this->*adv_)(loc): first candidate at or after loc; on failure the scan is over */
static int advance(scan_t *m, size_t loc)
{
  if (m->p->adv == UGX_ADV_NONE) /* advance_none, lib/matcher.cpp:957-960: no prediction, cur_ untouched */
    return 0;
  for (size_t k = loc; k < m->end; ++k)
  {
    if (cand(m, k))
    {
      set_current(m, k);
      return 1;
    }
  }
  set_current(m, m->end);
  return 0;
}

/* AbstractMatcher::skip('\n'), include/reflex/absmatcher.h:1198-1220 */
static int skip_newline(scan_t *m)
{
  const uint8_t *q = m->pos < m->end ? memchr(m->b + m->pos, '\n', m->end - m->pos) : NULL;
  if (q != NULL)
  {
    set_current(m, (size_t)(q - m->b) + 1);
    return 1;
  }
  set_current(m, m->end);
  return 0;
}

static inline int op_is_goto(uint32_t op) { return (op << 8) >= (op & 0xff000000u); }

/* evaluate one META code (include/reflex/pattern.h:930-952) after ch has been read; lib/matcher.cpp:272-404 */
static int meta_holds(const scan_t *m, unsigned meta, int ch, int bol)
{
  switch (meta)
  {
    case 0x0c: return ch == CH_EOF;                                                                /* EOB */
    case 0x0b: return 0;                                                                           /* BOB: got_ is never BOB after set_current */
    case 0x0a: return ch == CH_EOF || ch == '\n' || (ch == '\r' && peek_ch(m) == '\n');            /* EOL */
    case 0x09: return bol;                                                                         /* BOL */
    case 0x08: return at_we(m, ch, m->pos) && at_ew(m, ch);                                        /* EWE */
    case 0x07: return !at_we(m, ch, m->pos) && !at_ew(m, ch);                                      /* BWE */
    case 0x06: return !at_bw(m) && !at_wb(m);                                                      /* EWB */
    case 0x05: return at_bw(m) && at_wb(m);                                                        /* BWB */
    case 0x04: return at_we(m, ch, m->pos) != at_ew(m, ch);                                        /* NWE */
    case 0x03: return at_bw(m) != at_wb(m);                                                        /* NWB */
    case 0x02: return at_we(m, ch, m->pos) == at_ew(m, ch);                                        /* WBE */
    case 0x01: return at_bw(m) == at_wb(m);                                                        /* WBB */
    default: return 0;
  }
}

/*
 * One anchored attempt of the opcode interpreter at cur_ (lib/matcher.cpp:94-546).
 * Returns 1 when the ^-anchor fast path moved cur_ to the next line and the
 * caller must restart at `scan`, else 0 with cap/cur/pos set.
 */
static int run_dfa(scan_t *m, size_t *retry_io)
{
  const ora_pattern *p = m->p;
  const ugx_prefilter *pf = &p->pf;
  const uint32_t *opc = p->opc;
  const int W = (p->flags & UGX_OPT_W) != 0;
  size_t retry = *retry_io;
  {
    int ch = m->got;
    int bol = m->got == '\n';
    m->cap = 0;
    if (!W || at_wb(m))
    {
      if (pf->bol && !bol) /* :110-112 */
        if (skip_newline(m))
          return 1;
      uint32_t pc = 0;
      uint32_t back = NONE;
      size_t bpos = 0;
      for (;;)
      {
        uint32_t op = opc[pc];
        uint32_t jump;
        if (!op_is_goto(op))
        {
          if ((op >> 24) == 0xfe) /* TAKE at state entry, :139-150 */
          {
            if (!W || at_we(m, peek_ch(m), m->pos))
            {
              m->cap = op & 0xffffff;
              m->cur = m->pos;
            }
            ++pc;
            continue;
          }
          /* a block of META edges, :190-445 */
          if (ch == CH_EOF)
            break;
          ch = get_ch(m);
          int metas = 5;
          jump = NONE;
          for (;;)
          {
            if (jump == NONE || back == NONE)
            {
              if (!op_is_goto(op))
              {
                unsigned code = op >> 24;
                if (code == 0xfe) /* TAKE seen after ch was read: the match ends before ch, :207-217 */
                {
                  if (!W || at_we(m, ch, m->pos - 1))
                  {
                    m->cap = op & 0xffffff;
                    m->cur = m->pos;
                    if (ch != CH_EOF)
                      --m->cur;
                  }
                }
                else if (code != 0xff) /* 0xff: second word of a LONG jump, skipped */
                {
                  if (metas > 0 && jump == NONE && meta_holds(m, code, ch, bol))
                  {
                    --metas;
                    jump = op & 0xffff;
                    if (jump == IDX_LONG)
                      jump = opc[++pc] & 0xffffff;
                  }
                }
                op = opc[++pc];
                continue;
              }
              else if (ch != CH_EOF && op != OP_HALT)
              {
                if (jump == NONE)
                  break;
                if (back == NONE)
                {
                  back = pc;
                  bpos = m->pos - m->txt - 1;
                }
              }
            }
            if (jump == NONE)
            {
              if (back != NONE && bpos + 1 == m->pos - m->txt)
              {
                pc = back;
                op = opc[pc];
                back = NONE;
              }
              break;
            }
            if (back == pc)
              bpos = m->pos - m->txt - 1;
            pc = jump;
            op = opc[pc];
            jump = NONE;
          }
          if (ch == CH_EOF)
            break;
        }
        else
        {
          if (op == OP_HALT) /* :448-459 */
          {
            if (back != NONE)
            {
              m->pos = m->txt + bpos;
              pc = back;
              back = NONE;
              continue;
            }
            break;
          }
          if (ch == CH_EOF)
            break;
          ch = get_ch(m);
          if (ch == CH_EOF)
            break;
        }
        /* find the byte range that covers ch, :467-502 */
        while ((uint32_t)ch < (op >> 24) || (uint32_t)ch > ((op >> 16) & 0xff))
          op = opc[++pc];
        jump = op & 0xffff;
        if (jump == 0) /* back at the start state without an accept, :504-527 */
        {
          if (m->cap == 0)
          {
            if (m->cur + 1 == m->pos)
            {
              ++m->cur;
              if (retry > 0)
                --retry;
            }
            else
            {
              while (m->cur + 1 < m->pos && !(pf->fst[m->b[m->cur + 1] >> 3] >> (m->b[m->cur + 1] & 7) & 1))
              {
                ++m->cur;
                if (retry > 0)
                  --retry;
              }
            }
          }
        }
        else if (jump >= IDX_LONG)
        {
          if (jump == IDX_HALT)
          {
            if (back != NONE)
            {
              pc = back;
              m->pos = m->txt + bpos;
              back = NONE;
              continue;
            }
            break;
          }
          jump = opc[pc + 1] & 0xffffff;
        }
        pc = jump;
      }
    }
  }
  *retry_io = retry;
  return 0;
}

/*
 * One call of Matcher::match(Const::FIND), lib/matcher.cpp:42-750, on an
 * in-place buffer.  Returns the accept index (0 = no further match) and leaves
 * txt/len/cur/pos/got as the reference does.
 */
static uint32_t match_find(scan_t *m)
{
  const ora_pattern *p = m->p;
  const ugx_prefilter *pf = &p->pf;
  const int W = (p->flags & UGX_OPT_W) != 0;
  size_t retry = 0;
  m->len = 0;
  m->txt = m->cur;
  if (advance(m, m->cur)) /* :52 */
  {
    if (pf->lbk > 0) /* :54-70 */
    {
      size_t s = m->cur;
      if (s > m->txt)
      {
        size_t n = s - m->txt;
        if (pf->lbk != 0xffff && pf->lbk < n)
          n = pf->lbk;
        while (n-- > 0 && (pf->cbk[m->b[s - 1] >> 3] >> (m->b[s - 1] & 7) & 1))
        {
          --s;
          ++retry;
        }
        m->cur -= retry;
        retry = retry > pf->lbm ? retry - pf->lbm : 0;
      }
    }
    else if (pf->one) /* :71-83 */
    {
      size_t k = m->cur + pf->len;
      int ch = k < m->end ? m->b[k] : CH_EOF;
      if (!W || (at_wb(m) && (at_end(m) || at_we(m, ch, k))))
      {
        m->txt = m->cur;
        m->len = pf->len;
        set_current(m, k);
        return m->cap = 1;
      }
    }
  }
  else if (p->adv != UGX_ADV_NONE) /* a failed advance left cur_ at the end: nothing can match there */
  {
    m->txt = m->cur;
    return m->cap = 0;
  }
  set_current(m, m->cur);

scan:
  m->txt = m->cur;
  if (run_dfa(m, &retry))
    goto scan;
  if (m->cap == 0) /* :621-680 */
  {
    if (!at_end(m))
    {
      if (retry > 0)
      {
        --retry;
        set_current(m, m->cur + 1);
        goto scan;
      }
      if (m->cur < m->pos)
      {
        if (advance(m, m->cur + 1))
        {
          if (pf->lbk > 0)
          {
            size_t s = m->cur;
            if (s > m->txt + 1)
            {
              size_t n = s - m->txt - 1;
              if (pf->lbk != 0xffff && pf->lbk < n)
                n = pf->lbk;
              while (n-- > 0 && (pf->cbk[m->b[s - 1] >> 3] >> (m->b[s - 1] & 7) & 1))
              {
                --s;
                ++retry;
              }
              m->cur -= retry;
              retry = retry > pf->lbm ? retry - pf->lbm : 0;
            }
            set_current(m, m->cur);
            goto scan;
          }
          if (!pf->one)
            goto scan;
          size_t k = m->cur + pf->len;
          int ch = k < m->end ? m->b[k] : CH_EOF;
          if (W && (!at_wb(m) || !(at_end(m) || at_we(m, ch, k))))
            goto scan;
          m->txt = m->cur;
          m->len = pf->len;
          set_current(m, k);
          return m->cap = 1;
        }
        if (p->adv != UGX_ADV_NONE)
        {
          m->txt = m->cur;
          return m->cap = 0;
        }
      }
    }
    m->txt = m->cur;
  }
  m->len = m->cur - m->txt; /* :681 */
  if (m->len == 0)
  {
    m->pos = m->cur;
    if (at_end(m))
    {
      set_current(m, m->cur);
      return m->cap = 0;
    }
    if (m->cap != 0) /* an empty match, :692-721 */
    {
      if (p->flags & UGX_OPT_N) /* option N: accepted; the next find() starts one byte on */
      {
        set_current(m, m->cur + 1);
        return m->cap;
      }
      if (advance(m, m->cur + 1)) /* discarded: keep looking for a non-empty match */
        goto scan;
      return m->cap = 0;
    }
    set_current(m, m->cur + 1);
    goto scan;
  }
  set_current(m, m->cur);
  return m->cap;
}

/* ---- public ---- */

static int check_opcodes(const uint32_t *opc, uint32_t nop)
{
  for (uint32_t i = 0; i < nop; ++i)
  {
    uint32_t op = opc[i];
    if (op_is_goto(op))
      continue;
    unsigned code = op >> 24;
    if (code == 0xfd || code == 0xfc || code == 0xfb) /* REDO, TAIL, HEAD */
      return UGX_E_UNSUPPORTED;
    if (code >= 0x0d && code <= 0x0f) /* UND, IND, DED */
      return UGX_E_UNSUPPORTED;
  }
  return UGX_OK;
}

int ora_pattern_create(const uint32_t *opc, uint32_t nop, const ugx_prefilter *pf, uint32_t matcher_flags, ora_pattern **out)
{
  if (opc == NULL || nop == 0 || pf == NULL || out == NULL)
    return UGX_E_INVALID;
  int rc = check_opcodes(opc, nop);
  if (rc != UGX_OK)
    return rc;
  ora_pattern *p = calloc(1, sizeof(*p));
  if (p == NULL)
    return UGX_E_NOMEM;
  p->opc = malloc(sizeof(uint32_t) * (nop + 2));
  if (p->opc == NULL)
  {
    free(p);
    return UGX_E_NOMEM;
  }
  memcpy(p->opc, opc, sizeof(uint32_t) * nop);
  p->opc[nop] = p->opc[nop + 1] = OP_HALT;
  p->nop = nop;
  p->pf = *pf;
  p->flags = matcher_flags;
  p->adv = select_advance(pf, matcher_flags);
  if (pf->len == 0 && pf->pin >= 1 && pf->pin <= 16)
  {
    for (uint32_t i = 0; i < pf->pin; ++i)
    {
      p->pin_a[pf->chr[i]] = 1;
      p->pin_b[pf->chr[pf->pin + i]] = 1;
    }
  }
  *out = p;
  return UGX_OK;
}

int ora_pattern_load(const char *path, ora_pattern **out)
{
  FILE *f = fopen(path, "rb");
  if (f == NULL)
    return UGX_E_IO;
  ugx_file_header h;
  ugx_prefilter pf;
  int rc = UGX_E_IO;
  uint32_t *opc = NULL;
  if (fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, UGX_FILE_MAGIC, 8) == 0 &&
      h.prefilter_size == sizeof(pf) && fread(&pf, sizeof(pf), 1, f) == 1)
  {
    opc = malloc(sizeof(uint32_t) * (h.nop + 1));
    if (opc != NULL && fread(opc, sizeof(uint32_t), h.nop, f) == h.nop)
      rc = ora_pattern_create(opc, h.nop, &pf, h.matcher_flags, out);
  }
  free(opc);
  fclose(f);
  return rc;
}

void ora_pattern_destroy(ora_pattern *p)
{
  if (p != NULL)
  {
    free(p->opc);
    free(p);
  }
}

int ora_advance_kind(const ora_pattern *p) { return p->adv; }

static void scan_init(scan_t *m, const ora_pattern *p, const uint8_t *buf, uint64_t n)
{
  memset(m, 0, sizeof(*m));
  m->p = p;
  m->b = buf;
  m->end = (size_t)n;
  set_current(m, 0);
}

int ora_count_lines(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint64_t *count)
{
  scan_t m;
  scan_init(&m, p, buf, n);
  uint64_t c = 0;
  while (match_find(&m))
  {
    ++c;
    if (m.got != '\n') /* !at_bol(): src/ugrep.cpp:10583-10584 */
      skip_newline(&m);
  }
  *count = c;
  return UGX_OK;
}

int ora_count_matches(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint64_t *count)
{
  scan_t m;
  scan_init(&m, p, buf, n);
  uint64_t c = 0;
  while (match_find(&m))
    ++c;
  *count = c;
  return UGX_OK;
}

int ora_find_all(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint64_t base_offset, uint64_t base_line,
                 ugx_match *out, uint64_t cap, uint64_t *n_out)
{
  scan_t m;
  scan_init(&m, p, buf, n);
  uint64_t c = 0;
  uint64_t lno = 1;  /* lno_ */
  size_t lpb = 0;    /* lpb_ */
  uint32_t acc;
  while ((acc = match_find(&m)) != 0)
  {
    /* AbstractMatcher::lineno(), include/reflex/absmatcher.h:695-736 */
    for (; lpb < m.txt; ++lpb)
      lno += buf[lpb] == '\n';
    if (c < cap)
    {
      out[c].line = lno + base_line;
      out[c].offset = m.txt + base_offset; /* first(), absmatcher.h:901-905 */
      out[c].len = (uint32_t)m.len;
      out[c].cap = acc;
    }
    ++c;
  }
  *n_out = c;
  return c > cap ? UGX_E_OVERFLOW : UGX_OK;
}

/*
 * reflex::isutf8, lib/simd.cpp:169-421 (scalar form :396-418; the SSE2 / AVX2 / NEON loops, lib/simd_avx2.cpp:82-149,
 * apply the same rule 16 / 32 bytes at a time): no NUL, no byte C0 C1 F5..FF, a lead byte C2..DF / E0..EF / F0..F4 is
 * followed by exactly 1 / 2 / 3 continuation bytes 80..BF, no continuation byte anywhere else, no sequence cut off by
 * the end.  (Overlong 3- and 4-byte forms and surrogates pass: the reference calls its test "quick".)
 * ugrep uses it as its binary-file test: is_binary() = !isutf8(), src/ugrep.cpp:699-711.
 */
int ora_isutf8(const uint8_t *buf, uint64_t n)
{
  uint64_t i = 0;
  while (i < n)
  {
    int c = (int8_t)buf[i];
    if (c > 0)
    {
      ++i;
      continue;
    }
    ++i;
    if (c < -62 || c > -12 || i >= n || (buf[i++] & 0xc0) != 0x80)
      return 0;
    if (c >= -32 && (i >= n || (buf[i++] & 0xc0) != 0x80))
      return 0;
    if (c >= -16 && (i >= n || (buf[i++] & 0xc0) != 0x80))
      return 0;
  }
  return 1;
}

int ora_has_nul(const uint8_t *buf, uint64_t n) { return n > 0 && memchr(buf, 0, (size_t)n) != NULL; }

uint64_t ora_count_newlines(const uint8_t *buf, uint64_t n)
{
  uint64_t c = 0;
  for (uint64_t i = 0; i < n; ++i)
    c += buf[i] == '\n';
  return c;
}

int ora_candidates(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint8_t *bitmap)
{
  scan_t m;
  scan_init(&m, p, buf, n);
  memset(bitmap, 0, (size_t)((n + 7) / 8));
  for (size_t k = 0; k < m.end; ++k)
    if (cand(&m, k))
      bitmap[k >> 3] |= (uint8_t)(1u << (k & 7));
  return UGX_OK;
}

int ora_match_at(const ora_pattern *p, const uint8_t *buf, uint64_t n, uint64_t k, uint64_t *len)
{
  scan_t m;
  size_t retry = 0;
  scan_init(&m, p, buf, n);
  if (k > n)
    return 0;
  set_current(&m, (size_t)k);
  m.txt = m.cur;
  while (run_dfa(&m, &retry))
    m.txt = m.cur;
  *len = m.cap != 0 ? m.cur - m.txt : 0;
  return (int)m.cap;
}
