// literal_compile.cpp — ugx_compile_literal: the compiled form of ONE fixed string (`ugrep -F 'literal'`, config 1),
// produced here instead of by the reference's pattern compiler: the opcode words reflex::Pattern::encode_dfa writes for
// the chain DFA of a string (lib/pattern.cpp:2823-3063: per state the byte's GOTO, then the catch-all HALT; the last
// state is `TAKE 1`, HALT) and the prefilter fields its analysis leaves for a pure literal (lib/pattern.cpp:4286-4340:
// chr_/len_/one_, no predictor tables; :510-598: the two needle positions lcp_/lcs_ ranked by byte frequency, never
// Boyer-Moore on a SIMD build).  Byte-identical to `refscan dump -F -e LITERAL` (tests/test_literal_compile.py).
// The general regex / word-list compiler is NOT built.
#include <cstdlib>
#include <cstring>

#include "../../include/ugrep_b200.h"

namespace {

const unsigned char k_freq[256] = {
#include "byte_freq.inc"
};

int absdiff(int a, int b) { return a > b ? a - b : b - a; }

} // namespace

extern "C" int ugx_compile_literal(const uint8_t* lit, uint32_t len, uint32_t* opc, uint32_t cap, uint32_t* nop,
                                   ugx_prefilter* pf)
{
  if (lit == nullptr || nop == nullptr || pf == nullptr)
    return UGX_E_INVALID;
  // a pattern of 255 bytes or more is no longer "one string" to the reference (len_ is capped, one_ dropped); bytes the
  // command line cannot carry in one -F pattern: NUL, and the line breaks CNF::split cuts patterns at
  if (len == 0 || len > 254)
    return UGX_E_UNSUPPORTED;
  for (uint32_t i = 0; i < len; ++i)
    if (lit[i] == 0 || lit[i] == '\n' || lit[i] == '\r')
      return UGX_E_UNSUPPORTED;
  *nop = 2 * (len + 1);
  if (opc == nullptr || cap < *nop)
    return UGX_E_OVERFLOW;
  // state i sits at word 2 i: GOTO lit[i] -> state i + 1, then HALT for every other byte
  for (uint32_t i = 0; i < len; ++i)
  {
    opc[2 * i] = (static_cast<uint32_t>(lit[i]) << 24) | (static_cast<uint32_t>(lit[i]) << 16) | (2 * (i + 1));
    opc[2 * i + 1] = 0x00FFFFFFu;
  }
  opc[2 * len] = 0xFE000000u | 1u; // TAKE 1
  opc[2 * len + 1] = 0x00FFFFFFu;
  memset(pf, 0, sizeof(*pf));
  pf->len = len;
  pf->one = 1;
  memcpy(pf->chr, lit, len);
  memset(pf->bit, 0xff, sizeof(pf->bit));
  memset(pf->tap, 0xff, sizeof(pf->tap));
  memset(pf->pma, 0xff, sizeof(pf->pma));
  memset(pf->pmh, 0xff, sizeof(pf->pmh));
  pf->fst[lit[0] >> 3] |= static_cast<uint8_t>(1u << (lit[0] & 7));
  if (len > 1)
  {
    // the rarest byte is the first needle position, the next rarest the second (ties: the one farther from the first)
    int lcp = 0, lcs = 1;
    const int n = static_cast<int>(len);
    for (int i = 1; i < n; ++i)
    {
      const unsigned f = k_freq[lit[i]];
      if (k_freq[lit[lcp]] > f)
      {
        lcs = lcp;
        lcp = i;
      }
      else if (k_freq[lit[lcs]] > f || (k_freq[lit[lcs]] == f && absdiff(lcp, lcs) < absdiff(lcp, i)))
        lcs = i;
    }
    // adjacent positions are correlated: spread them apart
    if (n == 3 && (lcp == 1 || lcs == 1))
    {
      lcp = 0;
      lcs = 2;
    }
    else if (n > 3 && (lcp + 1 == lcs || lcs + 1 == lcp))
    {
      unsigned best = 255;
      for (int i = 0; i < n; ++i)
        if (i > lcp + 1 || i + 1 < lcp)
        {
          const unsigned f = k_freq[lit[i]];
          if (best > f)
          {
            lcs = i;
            best = f;
          }
        }
    }
    pf->lcp = static_cast<uint32_t>(lcp);
    pf->lcs = static_cast<uint32_t>(lcs);
  }
  return UGX_OK;
}
