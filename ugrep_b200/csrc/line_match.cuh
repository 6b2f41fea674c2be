// line_match.cuh — the reference's find loop (Matcher::match(FIND), lib/matcher.cpp:42-750) confined to
// one line, as device code: prefilter candidate -> look-back -> DFA attempt -> resume.  Shared by the scan kernels.
#pragma once

#include "device_pattern.cuh"

namespace ugx {

// where advance_to() gets candidates from: a tile's candidate bitmap in shared memory for positions inside
// the tile, the per-position predicate cand() outside it
struct CandMap {
  const uint32_t* bits; // bitmap of [base, base + nbits), or nullptr
  uint64_t base;
  uint32_t nbits;
};

constexpr int LINE_DONE = 2;

__device__ __forceinline__ bool advance_to(const Text& t, const DevPattern& P, const Tables& T, const CandMap& cm,
                                           Cursor& m, uint64_t loc, uint64_t last, bool need_got = true)
{
  // first candidate in [loc, last]; `last` is the line's '\n' (or the final byte of the buffer)
  uint64_t k = loc;
  if (cm.bits != nullptr && k >= cm.base && k < cm.base + cm.nbits)
  {
    uint32_t i = static_cast<uint32_t>(k - cm.base);
    uint32_t wi = i >> 5;
    const uint32_t nw = (cm.nbits + 31) >> 5;
    uint32_t word = cm.bits[wi] & (0xffffffffu << (i & 31));
    for (;;)
    {
      if (word != 0)
      {
        uint64_t hit = cm.base + (wi << 5) + (__ffs(word) - 1);
        if (hit > last || hit >= t.end)
          return false;
        set_current(t, m, hit, need_got);
        return true;
      }
      if (++wi >= nw)
        break;
      if (cm.base + (wi << 5) > last)
        return false;
      word = cm.bits[wi];
    }
    k = cm.base + cm.nbits;
  }
  for (; k <= last && k < t.end; ++k)
  {
    if (cand(t, P, T, k))
    {
      set_current(t, m, k, need_got);
      return true;
    }
  }
  return false;
}

// one anchored attempt over the dense table (patterns without META edges); lib/matcher.cpp:125-150, 446-546
__device__ __forceinline__ int run_dfa_table(const Text& t, const DevPattern& P, const Tables& T, Cursor& m, uint32_t& retry)
{
  const bool W = (P.flags & UGX_OPT_W) != 0;
  m.cap = 0;
  if (W && !at_wb(t, P, m))
    return 0;
  uint32_t state = 0;
  for (;;)
  {
    // states 1 .. first_acc - 1 neither accept nor halt (pattern_host.hpp numbering): no table read for them
    uint32_t acc = (state == 0 || state >= P.first_acc) ? __ldg(P.accept + state) : 0u;
    if ((acc & 0x7fffffffu) != 0 && (!W || at_we(t, P, peek_ch(t, m), m.pos)))
    {
      m.cap = acc & 0x7fffffffu;
      m.cur = m.pos;
    }
    if (acc & 0x80000000u) // state without outgoing edges: HALT before reading
      break;
    if (m.pos >= t.end)
      break;
    uint32_t ch = t.raw(m.pos++);
    uint32_t nxt = T.next[state * P.ncls + T.cls[ch]];
    if (nxt == D_DEAD)
      break;
    if (nxt == 0 && m.cap == 0) // back at the start state without an accept, lib/matcher.cpp:504-527
    {
      if (m.cur + 1 == m.pos)
      {
        ++m.cur;
        if (retry > 0)
          --retry;
      }
      else
      {
        while (m.cur + 1 < m.pos && !bit256(P.fst, t.raw(m.cur + 1)))
        {
          ++m.cur;
          if (retry > 0)
            --retry;
        }
      }
    }
    state = nxt;
  }
  return 0;
}

// one anchored attempt with the opcode interpreter (patterns with META edges); lib/matcher.cpp:94-546
__device__ inline int run_dfa_opc(const Text& t, const DevPattern& P, Cursor& m, uint32_t& retry)
{
  const bool W = (P.flags & UGX_OPT_W) != 0;
  const uint32_t* __restrict__ opc = P.opc;
  int ch = m.got;
  const bool bol = m.got == '\n';
  m.cap = 0;
  if (W && !at_wb(t, P, m))
    return 0;
  if (P.bol && !bol) // ^-anchored pattern away from a line start: the rest of this line cannot match
    return LINE_DONE;
  uint32_t pc = 0;
  uint32_t back = D_NONE;
  uint64_t bpos = 0;
  for (;;)
  {
    uint32_t op = __ldg(opc + pc);
    uint32_t jump;
    if (!d_op_is_goto(op))
    {
      if ((op >> 24) == 0xfe)
      {
        if (!W || at_we(t, P, peek_ch(t, m), m.pos))
        {
          m.cap = op & 0xffffff;
          m.cur = m.pos;
        }
        ++pc;
        continue;
      }
      if (ch == D_EOF)
        break;
      ch = get_ch(t, m);
      int metas = 5;
      jump = D_NONE;
      for (;;)
      {
        if (jump == D_NONE || back == D_NONE)
        {
          if (!d_op_is_goto(op))
          {
            uint32_t code = op >> 24;
            if (code == 0xfe)
            {
              if (!W || at_we(t, P, ch, m.pos - 1))
              {
                m.cap = op & 0xffffff;
                m.cur = m.pos;
                if (ch != D_EOF)
                  --m.cur;
              }
            }
            else if (code != 0xff)
            {
              if (metas > 0 && jump == D_NONE && meta_holds(t, P, m, code, ch, bol))
              {
                --metas;
                jump = op & 0xffff;
                if (jump == D_IDX_LONG)
                  jump = __ldg(opc + ++pc) & 0xffffff;
              }
            }
            op = __ldg(opc + ++pc);
            continue;
          }
          else if (ch != D_EOF && op != D_OP_HALT)
          {
            if (jump == D_NONE)
              break;
            if (back == D_NONE)
            {
              back = pc;
              bpos = m.pos - m.txt - 1;
            }
          }
        }
        if (jump == D_NONE)
        {
          if (back != D_NONE && bpos + 1 == m.pos - m.txt)
          {
            pc = back;
            op = __ldg(opc + pc);
            back = D_NONE;
          }
          break;
        }
        if (back == pc)
          bpos = m.pos - m.txt - 1;
        pc = jump;
        op = __ldg(opc + pc);
        jump = D_NONE;
      }
      if (ch == D_EOF)
        break;
    }
    else
    {
      if (op == D_OP_HALT)
      {
        if (back != D_NONE)
        {
          m.pos = m.txt + bpos;
          pc = back;
          back = D_NONE;
          continue;
        }
        break;
      }
      if (ch == D_EOF)
        break;
      ch = get_ch(t, m);
      if (ch == D_EOF)
        break;
    }
    while (static_cast<uint32_t>(ch) < (op >> 24) || static_cast<uint32_t>(ch) > ((op >> 16) & 0xff))
      op = __ldg(opc + ++pc);
    jump = op & 0xffff;
    if (jump == 0)
    {
      if (m.cap == 0)
      {
        if (m.cur + 1 == m.pos)
        {
          ++m.cur;
          if (retry > 0)
            --retry;
        }
        else
        {
          while (m.cur + 1 < m.pos && !bit256(P.fst, t.raw(m.cur + 1)))
          {
            ++m.cur;
            if (retry > 0)
              --retry;
          }
        }
      }
    }
    else if (jump >= D_IDX_LONG)
    {
      if (jump == D_IDX_HALT)
      {
        if (back != D_NONE)
        {
          pc = back;
          m.pos = m.txt + bpos;
          back = D_NONE;
          continue;
        }
        break;
      }
      jump = __ldg(opc + pc + 1) & 0xffffff;
    }
    pc = jump;
  }
  return 0;
}

__device__ __forceinline__ uint32_t look_back(const Text& t, const DevPattern& P, Cursor& m, uint64_t floor_pos)
{
  // walk back over cbk_ bytes from cur-1 down to floor_pos; lib/matcher.cpp:54-70, 639-654
  uint32_t retry = 0;
  uint64_t s = m.cur;
  if (s > floor_pos)
  {
    uint64_t n = s - floor_pos;
    if (P.lbk != 0xffff && P.lbk < n)
      n = P.lbk;
    while (n-- > 0 && bit256(P.cbk, t.raw(s - 1)))
    {
      --s;
      ++retry;
    }
    m.cur -= retry;
    retry = retry > P.lbm ? retry - P.lbm : 0;
  }
  return retry;
}

// one Matcher::match(FIND) confined to the line whose '\n' (or last byte) is at `last`.
// returns the accept index, or 0 when the line has no further match
template <bool HAS_META>
__device__ inline uint32_t find_in_line(const Text& t, const DevPattern& P, const Tables& T, const CandMap& cm, Cursor& m, uint64_t last)
{
  const bool W = (P.flags & UGX_OPT_W) != 0;
  const bool G = HAS_META || W; // got_ is needed
  uint32_t retry = 0;
  m.len = 0;
  m.txt = m.cur;
  if (!advance_to(t, P, T, cm, m, m.cur, last, G))
    return 0;
  if (P.lbk > 0)
  {
    retry = look_back(t, P, m, m.txt);
  }
  else if (P.one)
  {
    uint64_t k = m.cur + P.len;
    int ch = k < t.end ? static_cast<int>(t.raw(k)) : D_EOF;
    if (!W || (at_wb(t, P, m) && (m.pos >= t.end || at_we(t, P, ch, k))))
    {
      m.txt = m.cur;
      m.len = P.len;
      set_current(t, m, k, G);
      return m.cap = 1;
    }
  }
  set_current(t, m, m.cur, G);
  for (;;)
  {
    m.txt = m.cur;
    int r = HAS_META ? run_dfa_opc(t, P, m, retry) : run_dfa_table(t, P, T, m, retry);
    if (r == LINE_DONE)
      return 0;
    if (m.cap == 0)
    {
      if (m.pos < t.end)
      {
        if (retry > 0)
        {
          --retry;
          set_current(t, m, m.cur + 1, G);
          continue;
        }
        if (m.cur < m.pos)
        {
          if (!advance_to(t, P, T, cm, m, m.cur + 1, last, G))
            return 0;
          if (P.lbk > 0)
          {
            retry = look_back(t, P, m, m.txt + 1);
            set_current(t, m, m.cur, G);
            continue;
          }
          if (!P.one)
            continue;
          uint64_t k = m.cur + P.len;
          int ch = k < t.end ? static_cast<int>(t.raw(k)) : D_EOF;
          if (W && (!at_wb(t, P, m) || !(m.pos >= t.end || at_we(t, P, ch, k))))
            continue;
          m.txt = m.cur;
          m.len = P.len;
          set_current(t, m, k, G);
          return m.cap = 1;
        }
      }
      m.txt = m.cur;
    }
    m.len = static_cast<uint32_t>(m.cur - m.txt);
    if (m.len == 0)
    {
      m.pos = m.cur;
      if (m.pos >= t.end)
        return 0;
      if (m.cap != 0)
      {
        if (P.flags & UGX_OPT_N)
        {
          // option N: the empty match is reported, the next find() starts one byte on (lib/matcher.cpp:715-721)
          set_current(t, m, m.cur + 1, G);
          return m.cap;
        }
        if (!advance_to(t, P, T, cm, m, m.cur + 1, last, G))
          return 0;
        continue;
      }
      if (m.cur + 1 > last)
        return 0;
      set_current(t, m, m.cur + 1, G);
      continue;
    }
    set_current(t, m, m.cur, G);
    return m.cap;
  }
}

} // namespace ugx
