// wordlist_compile.cpp — ugx_compile_words: the compiled form of a list of fixed strings (`ugrep -F -f words.txt`,
// `ugrep -F -e A -e B`; config 2), produced here instead of by the reference's pattern compiler, byte for byte
// (tests/test_literal_compile.py compares with `refscan dump` on random lists).
//
// What the reference does for an all-literal alternation, restated (paths into /root/reference):
//   * the strings go into a tree DFA in the order given, states created in insertion order, the first string to end in
//     a state gives it its accept index (lib/pattern.cpp:798-866; no subset construction: :286-311);
//   * analyze_dfa (:3812-4385) walks the tree breadth-first up to the first accepting state on every path and weighs an
//     s-t "cut" (new start states deeper in the DFA plus a look-back).  The bookkeeping that DECIDES about a cut is
//     restated below; a list for which the reference does cut is refused (UGX_E_UNSUPPORTED) — what follows a cut
//     (:4102-4262) is not restated.  Without a cut: the common literal prefix chr_/len_/one_ (:4286-4340), min_ (gen_min,
//     :4387-4424), the bitap / hashed predictor tables (gen_predict_match, :4426-4639);
//   * Pattern::init (:331-598) then derives npy_ and the needle positions lcp_/lcs_ and needle bytes pin_/chr_ from
//     bit_ and the byte-frequency table, or — with a literal prefix — lcp_/lcs_ from the prefix;
//   * encode_dfa (:2823-3063) writes the opcode words: per state [TAKE], the byte edges by descending byte (ranges to
//     the same target merged by compact_dfa, :2764-2790), a catch-all HALT; two-word LONG jumps beyond 64K words.
// The general regex compiler is NOT built.
#include <algorithm>
#include <bitset>
#include <cstring>
#include <map>
#include <set>
#include <vector>

#include "../../include/ugrep_b200.h"

namespace {

const unsigned char k_freq[256] = {
#include "byte_freq.inc"
};

constexpr uint32_t IDX_LONG = 0xfffe, IDX_HALT = 0xffff, GMAX = 0xfeffff;

struct Node {
  std::map<uint32_t, uint32_t> edge; // byte -> node
  uint32_t accept = 0;
  uint32_t first = 0;                 // breadth-first depth + 2 during the analysis, word index while encoding
  uint32_t index = 0;
};

struct Ranges256 {
  std::bitset<256> b;
  void add(uint32_t c) { b.set(c); }
  uint32_t count() const { return static_cast<uint32_t>(b.count()); }
};

uint32_t hash3(uint32_t h, uint32_t c) { return ((h << 3) ^ c) & (UGX_HASH - 1); }

bool accepting_target(const std::vector<Node>& t, uint32_t s) { return t[s].accept > 0 || t[s].edge.empty(); }

// the part of analyze_dfa that decides whether the DFA is cut: returns true when the reference would cut
bool analysis_cuts(std::vector<Node>& t, ugx_prefilter& pf)
{
  bool searching = false;
  uint32_t fin_depth = 0xffff, fin_count = 0;
  uint32_t cut_depth = 0, cut_fin_count = 0, cut_span = 0, cut_count = 0xffff, min_count = 0xffff, max_count = 0;
  uint32_t best_cut_depth = 0, best_cut_fin_count = 0, best_cut_span = 0, best_cut_count = 0xffff, best_min_count = 0xffff;
  uint32_t max_freq = 0;
  std::set<uint32_t> states{0}, next_states;
  t[0].first = 1;
  uint32_t count = 0, prev_min_count = 0xffff;
  for (uint32_t depth = 0; depth < 256; ++depth)
  {
    next_states.clear();
    Ranges256 next_chars;
    const bool is_more = fin_count == 0;
    for (uint32_t s : states)
      for (const auto& e : t[s].edge)
      {
        const uint32_t c = e.first, nx = e.second;
        if (depth == 0)
          pf.fst[c >> 3] |= static_cast<uint8_t>(1u << (c & 7));
        if (accepting_target(t, nx))
        {
          t[nx].first = 0;
          if (fin_depth == 0xffff)
            fin_depth = depth;
          ++fin_count;
          continue;
        }
        if (t[nx].first == 0 || t[nx].first > cut_depth + 1)
          next_chars.add(c);
        if (t[nx].first == 0)
          t[nx].first = depth + 2;
        next_states.insert(nx); // (a tree has no edge back to an earlier state)
      }
    count = next_chars.count();
    for (uint32_t c = 0; c < 256; ++c)
      if (next_chars.b[c] && k_freq[c] > max_freq)
        max_freq = k_freq[c];
    prev_min_count = min_count;
    if (count > max_count)
      max_count = count;
    if (count + fin_count < min_count)
      min_count = count + fin_count;
    if (is_more)
      cut_span = depth - cut_depth;
    if (searching)
    {
      bool make_cut;
      if (fin_count == 0)
        make_cut = cut_span > 6 && prev_min_count < 0xffff && prev_min_count > 8 && prev_min_count >= min_count;
      else
        make_cut = cut_span > 7 && prev_min_count < 0xffff && prev_min_count > 8 && min_count <= 8;
      if (make_cut)
      {
        const bool better = cut_span <= 2 ? cut_span > best_cut_span : (best_min_count >= prev_min_count && cut_span >= best_cut_span);
        if (better)
        {
          best_cut_count = cut_count;
          best_cut_depth = cut_depth;
          best_cut_fin_count = cut_fin_count;
          best_cut_span = cut_span;
          best_min_count = prev_min_count;
          searching = false;
        }
      }
    }
    if (!searching)
    {
      // (the recount at depth > 0 leaves out self-edges only: a tree has none, the count stands)
      cut_count = count + fin_count;
      cut_depth = depth;
      cut_fin_count = fin_count;
      max_freq = 0;
      max_count = count;
      min_count = cut_count;
      searching = true;
    }
    states.swap(next_states);
    if (count <= fin_count || (!is_more && cut_span < 2))
    {
      if (is_more)
        ++cut_span;
      if (min_count < cut_count && min_count < best_min_count)
        if (cut_span >= 2 && prev_min_count < 0xffff && prev_min_count >= 64 && min_count <= 8)
        {
          best_cut_count = count + fin_count;
          best_cut_depth = depth;
          best_cut_fin_count = fin_count;
          best_cut_span = cut_span;
          best_min_count = min_count;
        }
      break;
    }
  }
  if (best_cut_depth > 0 || best_cut_span > 0)
  {
    bool better = false;
    if ((best_cut_span == 1 || min_count < best_min_count || best_cut_fin_count == cut_fin_count) && cut_count <= best_cut_count &&
        min_count <= best_min_count)
    {
      if (cut_span == 2 && fin_count > cut_count)
        better = min_count < best_min_count;
      else if (cut_span > best_cut_span)
        better = cut_fin_count == 0 || min_count < best_min_count;
      else if (cut_span >= 2 || cut_span == best_cut_span)
        better = min_count < best_min_count;
    }
    if (!better)
      cut_depth = best_cut_depth;
  }
  return cut_depth > 0;
}

struct Level {
  std::bitset<UGX_HASH> hashes;
  std::bitset<256> chars;
};

// gen_min (lib/pattern.cpp:4387-4424) for a tree without a cut
uint32_t gen_min(const std::vector<Node>& t, uint32_t start)
{
  uint32_t min = 8;
  std::set<uint32_t> prev, next{start};
  for (uint32_t level = 0; level < min; ++level)
  {
    bool none = true;
    prev.clear();
    prev.swap(next);
    for (uint32_t from : prev)
    {
      const bool from_accepts = accepting_target(t, from);
      if (!from_accepts)
        for (const auto& e : t[from].edge)
        {
          none = false;
          if (min == level + 1)
            continue;
          if (accepting_target(t, e.second))
            min = level + 1;
          else
            next.insert(e.second);
        }
      if (from_accepts)
      {
        none = true;
        break;
      }
    }
    if (none)
      min = level;
  }
  return min;
}

void tap_accepting(ugx_prefilter& pf, uint32_t c, uint8_t mask)
{
  for (uint32_t h = c & 63u; h < UGX_BTAP; h += 64)
    pf.tap[h] &= mask;
}

void tap_pairs(ugx_prefilter& pf, const std::vector<Node>& t, uint32_t c, uint32_t next_state, uint8_t mask)
{
  for (const auto& e : t[next_state].edge)
    pf.tap[(c ^ (e.first << 6)) & (UGX_BTAP - 1)] &= mask;
}

// gen_predict_match (lib/pattern.cpp:4426-4639) for a tree without a cut, from the single start state `start`
void gen_predict(const std::vector<Node>& t, uint32_t start, ugx_prefilter& pf)
{
  const uint32_t min = gen_min(t, start);
  pf.min = min;
  const uint32_t levels = min > 4 ? min : 4;
  std::map<uint32_t, Level> cur, nxt;
  // level 0: the hash of one byte is the byte
  for (const auto& e : t[start].edge)
  {
    const uint32_t c = e.first, ns = e.second;
    const bool acc = accepting_target(t, ns);
    Level& L = cur[ns];
    L.hashes.set(c);
    L.chars.set(c);
    pf.bit[c] &= static_cast<uint8_t>(~1u);
    pf.pmh[c] &= static_cast<uint8_t>(~1u);
    pf.pma[c] &= static_cast<uint8_t>(acc ? ~0xc0u : ~0x40u);
    if (min <= 1)
    {
      if (acc)
        tap_accepting(pf, c, static_cast<uint8_t>(~1u));
      else
        tap_pairs(pf, t, c, ns, static_cast<uint8_t>(~1u));
    }
  }
  for (uint32_t level = 1; level < levels && !cur.empty(); ++level)
  {
    nxt.clear();
    const bool pass_on = level + 1 < levels;
    for (const auto& from : cur)
      for (const auto& e : t[from.first].edge)
      {
        const uint32_t c = e.first, ns = e.second;
        const bool acc = accepting_target(t, ns);
        Level* nl = pass_on ? &nxt[ns] : nullptr;
        if (level < min)
        {
          const uint8_t mask = static_cast<uint8_t>(~(1u << level));
          pf.bit[c] &= mask;
          // the pair (previous byte, this byte) at the previous level
          const uint8_t prev_mask = static_cast<uint8_t>(mask >> 1);
          for (uint32_t pc = 0; pc < 256; ++pc)
            if (from.second.chars[pc])
              pf.tap[(pc ^ (c << 6)) & (UGX_BTAP - 1)] &= prev_mask;
          if (level + 1 < min && nl != nullptr)
            nl->chars.set(c);
          else if (acc)
            tap_accepting(pf, c, mask);
          else
            tap_pairs(pf, t, c, ns, mask);
        }
        if (level < 4)
        {
          const uint8_t pmh_mask = static_cast<uint8_t>(~(1u << level));
          uint8_t pma_mask = static_cast<uint8_t>(~(1u << (6 - 2 * level)));
          if (level == 3 || acc)
            pma_mask &= static_cast<uint8_t>(~(1u << (7 - 2 * level)));
          for (uint32_t ph = 0; ph < UGX_HASH; ++ph)
            if (from.second.hashes[ph])
            {
              const uint32_t h = hash3(ph, c);
              pf.pmh[h] &= pmh_mask;
              pf.pma[h] &= pma_mask;
              if (nl != nullptr)
                nl->hashes.set(h);
            }
        }
        else if (level < min)
        {
          const uint8_t pmh_mask = static_cast<uint8_t>(~(1u << level));
          for (uint32_t ph = 0; ph < UGX_HASH; ++ph)
            if (from.second.hashes[ph])
            {
              const uint32_t h = hash3(ph, c);
              pf.pmh[h] &= pmh_mask;
              if (nl != nullptr)
                nl->hashes.set(h);
            }
        }
      }
    cur.swap(nxt);
  }
}

int absdiff(int a, int b) { return a > b ? a - b : b - a; }

// Pattern::init after the analysis (lib/pattern.cpp:331-598)
void post_analysis(ugx_prefilter& pf)
{
  if (pf.len == 0)
  {
    const uint32_t min_ = pf.min;
    if (min_ > 0)
    {
      if (min_ < 8)
      {
        const uint8_t mask = static_cast<uint8_t>(~((1u << min_) - 1));
        for (uint32_t i = 0; i < 256; ++i)
          pf.bit[i] |= mask;
        for (uint32_t i = 0; i < UGX_BTAP; ++i)
          pf.tap[i] |= mask;
      }
      uint32_t npy = 0;
      for (uint32_t i = 0; i < 256; ++i)
      {
        pf.bit[i] |= static_cast<uint8_t>(~((1u << min_) - 1));
        for (uint32_t b = 0; b < 8; ++b)
          npy += ((pf.bit[i] >> b) & 1u) == 0;
      }
      pf.npy = npy / min_;
    }
    // needle positions: per position k < min the bytes that can stand there; few and rare bytes make a good needle
    const uint32_t pinmax = 16, freqmax1 = 20, freqmax2 = 251, freqmax3 = 300;
    uint32_t nlcp = 65535, nlcs = 65535, freqlcp = 255, freqlcs = 255;
    int lcp = 0, lcs = 0;
    const uint32_t min = min_ > 1 ? min_ : 1;
    uint8_t score[9][3];
    size_t scores = 0;
    for (uint32_t k = 0; k < min; ++k)
    {
      const uint8_t mask = static_cast<uint8_t>(1u << k);
      uint32_t n = 0, max = 0, sum = 0;
      for (uint32_t i = 0; i < 256 && n <= pinmax; ++i)
        if ((pf.bit[i] & mask) == 0)
        {
          ++n;
          const uint32_t f = k_freq[i];
          if (f > max)
            max = f;
          sum += f;
        }
      if (n > 0 && n <= pinmax && max <= freqmax2)
      {
        const uint32_t mm = std::min<uint32_t>((sum + n - 1) / n * ((n > 8) + 1), 255);
        const uint8_t m = static_cast<uint8_t>(mm);
        if (m <= freqmax2)
        {
          size_t i;
          for (i = 0; i < scores; ++i)
            if (score[i][0] > m || (score[i][0] == m && score[i][2] > n))
            {
              memmove(score[i + 1], score[i], (scores - i) * 3);
              break;
            }
          score[i][0] = m;
          score[i][1] = static_cast<uint8_t>(k);
          score[i][2] = static_cast<uint8_t>(n);
          ++scores;
        }
      }
    }
    if (scores == 1 && min_ <= 3)
    {
      freqlcp = freqlcs = score[0][0];
      lcp = lcs = score[0][1];
      nlcp = nlcs = score[0][2];
      const uint32_t freqmax = (min_ > 1 || nlcp > 5) ? freqmax1 : freqmax2;
      if (freqlcp > freqmax)
        freqlcp = freqlcs = 255;
    }
    else if (scores >= 2)
    {
      freqlcp = score[0][0];
      lcp = score[0][1];
      nlcp = score[0][2];
      freqlcs = score[1][0];
      lcs = score[1][1];
      nlcs = score[1][2];
      if (lcp + 1 == lcs || lcs + 1 == lcp || (nlcp <= 8 && nlcs > 8))
        for (size_t i = 2; i < scores; ++i)
          if (score[i][2] <= 8 && absdiff(lcp, score[i][1]) > 1)
          {
            freqlcs = score[i][0];
            lcs = score[i][1];
            nlcs = score[i][2];
            break;
          }
    }
    uint32_t n = std::max(nlcp, nlcs);
    uint32_t freqmax = 2 * freqmax2;
    if (n > 8 && min_ >= 3)
      freqmax = freqmax3;
    pf.lcp = static_cast<uint32_t>(lcp);
    pf.lcs = static_cast<uint32_t>(lcs);
    if (n > 0 && n <= pinmax && freqlcp + freqlcs <= freqmax)
    {
      if (n > 8)
        n = 16;
      uint32_t j = 0, k = n;
      const uint8_t masklcp = static_cast<uint8_t>(1u << lcp), masklcs = static_cast<uint8_t>(1u << lcs);
      for (uint32_t i = 0; i < 256; ++i)
      {
        if ((pf.bit[i] & masklcp) == 0)
          pf.chr[j++] = static_cast<uint8_t>(i);
        if ((pf.bit[i] & masklcs) == 0)
          pf.chr[k++] = static_cast<uint8_t>(i);
      }
      for (; j < n; ++j)
        pf.chr[j] = pf.chr[j - 1];
      for (; k < 2 * n; ++k)
        pf.chr[k] = pf.chr[k - 1];
      pf.pin = n;
    }
  }
  else if (pf.len > 1)
  {
    // a literal prefix: the rarest byte is the first needle position, the next rarest the second
    int lcp = 0, lcs = 1;
    const int n = static_cast<int>(pf.len);
    for (int i = 1; i < n; ++i)
    {
      const unsigned f = k_freq[pf.chr[i]];
      if (k_freq[pf.chr[lcp]] > f)
      {
        lcs = lcp;
        lcp = i;
      }
      else if (k_freq[pf.chr[lcs]] > f || (k_freq[pf.chr[lcs]] == f && absdiff(lcp, lcs) < absdiff(lcp, i)))
        lcs = i;
    }
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
          const unsigned f = k_freq[pf.chr[i]];
          if (best > f)
          {
            lcs = i;
            best = f;
          }
        }
    }
    pf.lcp = static_cast<uint32_t>(lcp);
    pf.lcs = static_cast<uint32_t>(lcs);
  }
}

struct OutEdge {
  uint32_t lo, hi, target; // target = node, or 0xffffffff for the dead state
};

// compact_dfa + encode_dfa (lib/pattern.cpp:2764-3063)
int encode(std::vector<Node>& t, std::vector<uint32_t>& opc)
{
  const uint32_t DEAD = 0xffffffffu;
  std::vector<std::vector<OutEdge>> edges(t.size());
  uint32_t nop = 0;
  for (size_t s = 0; s < t.size(); ++s)
  {
    // ranges of adjacent bytes to the same target are merged
    std::vector<OutEdge>& out = edges[s];
    for (const auto& e : t[s].edge)
    {
      if (!out.empty() && out.back().target == e.second && out.back().hi + 1 == e.first)
        out.back().hi = e.first;
      else
        out.push_back(OutEdge{e.first, e.first, e.second});
    }
    // the dead state takes what the edges leave: from the first uncovered byte up to 0xff (the interpreter searches the
    // edges from the highest byte down, so the edges above it are found first)
    uint32_t hi = 0;
    for (const OutEdge& e : out)
      if (e.lo == hi)
        hi = e.hi + 1;
    t[s].first = t[s].index = nop;
    nop += static_cast<uint32_t>(out.size());
    if (hi <= 0xff)
    {
      out.insert(std::upper_bound(out.begin(), out.end(), hi, [](uint32_t v, const OutEdge& e) { return v < e.lo; }),
                 OutEdge{hi, 0xff, DEAD});
      ++nop;
    }
    nop += t[s].accept > 0;
    if (nop > GMAX)
      return UGX_E_UNSUPPORTED;
  }
  const bool wide = nop > IDX_LONG;
  auto is_long = [&](size_t s, uint32_t target) {
    return target != DEAD && ((t[target].first > t[s].first && t[target].first >= IDX_LONG / 2) || t[target].index >= IDX_LONG);
  };
  if (wide)
  {
    // over 64K words: jumps far ahead or far back take two words; the states move, the first pass's positions decide
    nop = 0;
    for (size_t s = 0; s < t.size(); ++s)
    {
      t[s].index = nop;
      for (const OutEdge& e : edges[s])
        nop += is_long(s, e.target) ? 2 : 1;
      nop += t[s].accept > 0;
      if (nop > GMAX)
        return UGX_E_UNSUPPORTED;
    }
  }
  opc.clear();
  opc.reserve(nop);
  for (size_t s = 0; s < t.size(); ++s)
  {
    if (t[s].accept > 0)
      opc.push_back(0xfe000000u | (t[s].accept & 0xffffffu));
    for (auto e = edges[s].rbegin(); e != edges[s].rend(); ++e)
    {
      const uint32_t head = (e->lo << 24) | (e->hi << 16);
      if (e->target == DEAD)
        opc.push_back(head | IDX_HALT);
      else if (wide && is_long(s, e->target))
      {
        opc.push_back(head | IDX_LONG);
        opc.push_back(0xff000000u | (t[e->target].index & 0xffffffu));
      }
      else
        opc.push_back(head | t[e->target].index);
    }
  }
  return opc.size() == nop ? UGX_OK : UGX_E_INVALID;
}

} // namespace

extern "C" int ugx_compile_words(const uint8_t* const* words, const uint32_t* lens, uint32_t nwords, uint32_t* opc, uint32_t cap,
                                 uint32_t* nop, ugx_prefilter* pf)
{
  return ugx_compile_words_ex(words, lens, nwords, 0, opc, cap, nop, pf);
}

extern "C" int ugx_compile_words_ex(const uint8_t* const* words, const uint32_t* lens, uint32_t nwords, uint32_t options,
                                    uint32_t* opc, uint32_t cap, uint32_t* nop, ugx_prefilter* pf)
{
  if (words == nullptr || lens == nullptr || nop == nullptr || pf == nullptr || nwords == 0 || (options & ~1u) != 0)
    return UGX_E_INVALID;
  const bool icase = (options & UGX_COMPILE_ICASE) != 0;
  try
  {
    // ---- the tree, in the order given
    std::vector<Node> t(1);
    for (uint32_t w = 0; w < nwords; ++w)
    {
      if (lens[w] == 0)
        return UGX_E_UNSUPPORTED; // an empty string matches everything: ugrep drops it before the compiler sees it
      uint32_t r = 0;
      for (uint32_t i = 0; i < lens[w]; ++i)
      {
        uint32_t c = words[w][i];
        if (c == 0 || c == '\n' || c == '\r')
          return UGX_E_UNSUPPORTED;
        if (icase && c >= 'A' && c <= 'Z')
          c += 'a' - 'A'; // (lib/pattern.cpp:834)
        auto it = t[r].edge.find(c);
        if (it == t[r].edge.end())
        {
          const uint32_t fresh = static_cast<uint32_t>(t.size());
          t[r].edge[c] = fresh;
          t.emplace_back();
          r = fresh;
        }
        else
          r = it->second;
      }
      if (t[r].accept == 0)
        t[r].accept = w + 1;
    }
    if (icase)
      for (Node& s : t) // an uppercase twin for every edge on a lowercase letter (lib/pattern.cpp:292-309)
        for (uint32_t c = 'a'; c <= 'z'; ++c)
        {
          auto it = s.edge.find(c);
          if (it != s.edge.end())
            s.edge[c - ('a' - 'A')] = it->second;
        }
    memset(pf, 0, sizeof(*pf));
    // ---- analysis
    if (analysis_cuts(t, *pf))
      return UGX_E_UNSUPPORTED; // the reference would cut this DFA and search with a look-back: not restated
    uint32_t state = 0;
    bool one = true;
    while (t[state].accept == 0)
    {
      if (t[state].edge.size() != 1)
      {
        one = false;
        break;
      }
      if (pf->len >= 255)
      {
        one = false;
        break;
      }
      pf->chr[pf->len++] = static_cast<uint8_t>(t[state].edge.begin()->first);
      state = t[state].edge.begin()->second;
    }
    if (pf->len == 1 && t[state].accept == 0 && !t[state].edge.empty())
    {
      // a one-byte prefix is only kept when the pattern can end there
      pf->len = 0; // (chr_[0] keeps the byte, as in the reference)
      one = false;
      state = 0;
    }
    if (t[state].accept > 0 && !t[state].edge.empty())
      one = false;
    pf->one = one ? 1u : 0u;
    memset(pf->bit, 0xff, sizeof(pf->bit));
    memset(pf->tap, 0xff, sizeof(pf->tap));
    memset(pf->pma, 0xff, sizeof(pf->pma));
    memset(pf->pmh, 0xff, sizeof(pf->pmh));
    if (pf->len == 0 || t[state].accept == 0)
      gen_predict(t, state, *pf);
    post_analysis(*pf);
    // ---- opcode words
    std::vector<uint32_t> words_out;
    const int rc = encode(t, words_out);
    if (rc != UGX_OK)
      return rc;
    *nop = static_cast<uint32_t>(words_out.size());
    if (opc == nullptr || cap < *nop)
      return UGX_E_OVERFLOW;
    memcpy(opc, words_out.data(), words_out.size() * 4);
    return UGX_OK;
  }
  catch (const std::bad_alloc&)
  {
    return UGX_E_NOMEM;
  }
}

// a regex without operators other than top-level alternation and escapes of single characters: the strings it stands for
extern "C" int ugx_compile_plain_regex(const uint8_t* regex, uint32_t len, uint32_t options, uint32_t* opc, uint32_t cap,
                                       uint32_t* nop, ugx_prefilter* pf)
{
  if (regex == nullptr || nop == nullptr || pf == nullptr)
    return UGX_E_INVALID;
  try
  {
    std::vector<std::vector<uint8_t>> alts(1);
    bool quoted = false;
    uint32_t from = 0;
    if (len >= 4 && memcmp(regex, "(?i)", 4) == 0) // the inline form of -i (lib/pattern.cpp:620-752)
    {
      options |= UGX_COMPILE_ICASE;
      from = 4;
    }
    for (uint32_t i = from; i < len; ++i)
    {
      const uint8_t c = regex[i];
      if (quoted)
      {
        if (c == '\\' && i + 1 < len && regex[i + 1] == 'E')
        {
          quoted = false;
          ++i;
        }
        else
          alts.back().push_back(c);
        continue;
      }
      if (c == '|')
      {
        alts.emplace_back();
        continue;
      }
      if (c == '\\')
      {
        if (i + 1 >= len)
          return UGX_E_UNSUPPORTED;
        const uint8_t e = regex[++i];
        if (e == 'Q')
          quoted = true;
        else if (e == 't')
          alts.back().push_back('\t');
        else if (e == 'f')
          alts.back().push_back('\f');
        else if (e == 'v')
          alts.back().push_back('\v');
        else if (e == 'a')
          alts.back().push_back('\a');
        else if (e != 0 && strchr("\\.[](){}*+?|^$!\"#%&',-/:;@`", e) != nullptr)
          alts.back().push_back(e); // an escaped operator / punctuation character stands for itself
        else
          return UGX_E_UNSUPPORTED; // \d \w \s \b \< \p{..} \1 ...: classes, anchors, back-references — and the escapes
                                    // the reference's converter rewrites (\~ \e \xHH ...), which leave its tree path
        continue;
      }
      if (c == '.' || c == '[' || c == ']' || c == '(' || c == ')' || c == '{' || c == '}' || c == '*' || c == '+' || c == '?' ||
          c == '^' || c == '$')
        return UGX_E_UNSUPPORTED; // (a bare ] or } is an ordinary character to the reference's parser in most places, an
                                  // error in some: refused, write \\] \\})
      if (c >= 0x80 && (options & UGX_COMPILE_ICASE) != 0)
        return UGX_E_UNSUPPORTED; // the reference's converter folds the case of non-ASCII letters into classes (not so
                                  // inside \\Q..\\E, which is how -F -i passes them)
      alts.back().push_back(c);
    }
    if (quoted)
      return UGX_E_UNSUPPORTED;
    std::vector<const uint8_t*> ptrs;
    std::vector<uint32_t> lens;
    for (const auto& a : alts)
    {
      if (a.empty())
        return UGX_E_UNSUPPORTED; // an empty alternative matches the empty string
      ptrs.push_back(a.data());
      lens.push_back(static_cast<uint32_t>(a.size()));
    }
    return ugx_compile_words_ex(ptrs.data(), lens.data(), static_cast<uint32_t>(ptrs.size()), options, opc, cap, nop, pf);
  }
  catch (const std::bad_alloc&)
  {
    return UGX_E_NOMEM;
  }
}
