// integration/b200matcher.hpp — the reference-side binding of the B200 scan path: a reflex::Matcher whose
// match(FIND) is served by libugrep_b200.so.
//
// This file is compiled INTO the reference (GerHobbelt/ugrep 7.4.2), not into this repository's library: it derives
// from the reference's own reflex::Matcher (include/reflex/matcher.h:47) and replaces the one virtual ugrep's search
// loops reach on this path — AbstractMatcher::match(Method), include/reflex/absmatcher.h:1413 — keeping every other
// member (find(), lineno(), first(), begin(), size(), bol(), eol(), skip(), at_bol(), ...; absmatcher.h:168-1220)
// coherent by setting the protected fields they read (absmatcher.h:1632-1660).  integration/Makefile builds the
// reference CLI with `new reflex::Matcher(...)` at src/ugrep.cpp:8902 turned into `new B200Matcher(...)` (a sed at build
// time; nothing of the reference is copied here).  It reads reflex::Pattern's compiled tables, which are protected:
// the translation unit is built with -fno-access-control (a maintainer would add `friend class B200Matcher;`).
//
// How a file is searched:
//   * the first find() on a new input makes sure the whole input is in memory (mmap'd files already are:
//     AbstractMatcher::buffer(base, size), absmatcher.h:542-591; streamed inputs are read to their end into the
//     matcher's own buffer, as peek_more() would block by block) and hands it to ugx_find_all_device: ONE device scan;
//   * every find() then replays the next record — txt_/len_/cap_/cur_/pos_/got_ as Matcher::match leaves them, and
//     lno_/lpb_/bol_ advanced the way lineno() would have (absmatcher.h:695-736) without re-reading the text;
//   * after skip('\n') (the only way ugrep moves the cursor itself, src/ugrep.cpp:3991, 10584, ...) records before
//     the cursor are dropped: at a line start the chain of matches is the same with or without the skipped ones.
// Patterns outside the library's scope (ugx_pattern_create returns UGX_E_UNSUPPORTED: lookahead, '\n' in the
// pattern, ...) and the methods SCAN / SPLIT / MATCH stay with the base class — a choice made on the reference side;
// the library itself never scans on the CPU.  UGREP_B200_REQUIRE=1 turns that into a hard error, UGREP_B200_VERBOSE=1
// reports which engine serves the pattern.
#ifndef UGREP_B200_MATCHER_HPP
#define UGREP_B200_MATCHER_HPP

#include <reflex/matcher.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ugrep_b200.h"

class B200Matcher : public reflex::Matcher {
 public:
  B200Matcher(const reflex::Pattern& pattern, const reflex::Input& input = reflex::Input(), const char *opt = NULL)
    : reflex::Matcher(pattern, input, opt)
  {
    upload();
    fresh_ = true;
  }
  B200Matcher(const B200Matcher& matcher) : reflex::Matcher(matcher), shared_(matcher.shared_)
  {
    if (shared_ && shared_->pattern != NULL)
      new_scanner();
    fresh_ = true;
  }
  virtual ~B200Matcher()
  {
    if (scanner_ != NULL)
      ugx_scanner_destroy(scanner_);
  }
  // one matcher per worker thread, the compiled pattern shared (src/ugrep.cpp:4204-4215)
  virtual reflex::Matcher *clone()
  {
    return new B200Matcher(*this);
  }
  virtual void reset(const char *opt = NULL)
  {
    reflex::Matcher::reset(opt);
    fresh_ = true;
  }

 protected:
  virtual size_t match(Method method)
  {
    if (scanner_ == NULL || method != Const::FIND)
      return reflex::Matcher::match(method);
    reset_text();
    if (fresh_ || got_ == Const::BOB) // reset() / input(), or an in-place buffer(base, size) since the last find()
      scan_input();
    for (;;)
    {
      if (next_ >= count_)
      {
        // no further match: the matcher is left at the end of the input (lib/matcher.cpp:672-690)
        txt_ = buf_ + end_;
        len_ = 0;
        cap_ = 0;
        set_current(end_);
        got_ = Const::EOB;
        return 0;
      }
      const ugx_match& m = record(next_);
      if (m.offset >= cur_)
        break;
      ++next_; // the caller moved past it with skip('\n')
    }
    const ugx_match& m = record(next_++);
    txt_ = buf_ + m.offset;
    len_ = m.len;
    cap_ = m.cap;
    // lineno() bookkeeping (absmatcher.h:695-736) from the record's line number instead of a pass over the text
    if (lpb_ < txt_)
    {
      if (m.line > lpb_line_)
      {
        lno_ += cml_ ? 1 : static_cast<size_t>(m.line - lpb_line_);
        const char *b = txt_;
        const void *nl = b > buf_ ? memrchr(buf_, '\n', b - buf_) : NULL;
        bol_ = nl != NULL ? static_cast<char*>(const_cast<void*>(nl)) + 1 : buf_;
        cpb_ = bol_;
        cno_ = 0;
      }
      lpb_ = txt_;
      lpb_line_ = m.line;
    }
    // an empty match (option N) leaves the cursor one byte on (lib/matcher.cpp:715-721)
    set_current(m.offset + (m.len > 0 ? m.len : 1));
    return cap_;
  }

 private:
  struct Shared {
    ugx_pattern *pattern;
    Shared() : pattern(NULL) { }
    ~Shared()
    {
      if (pattern != NULL)
        ugx_pattern_destroy(pattern);
    }
  };

  static bool env_set(const char *name)
  {
    const char *v = getenv(name);
    return v != NULL && *v != '\0' && *v != '0';
  }

  // hand reflex::Pattern's compiled form to the library: opcode words + the prefilter members Matcher reads
  // (include/reflex/pattern.h:1288-1334)
  void upload()
  {
    const reflex::Pattern& pat = *pat_;
    shared_ = std::make_shared<Shared>();
    if (pat.opc_ == NULL || pat.nop_ == 0)
      return unsupported("the pattern has no opcode table");
    ugx_prefilter pf;
    memset(&pf, 0, sizeof(pf));
    pf.len = static_cast<uint32_t>(pat.len_);
    pf.min = static_cast<uint32_t>(pat.min_);
    pf.pin = static_cast<uint32_t>(pat.pin_);
    pf.lcp = pat.lcp_;
    pf.lcs = pat.lcs_;
    pf.bmd = static_cast<uint32_t>(pat.bmd_);
    pf.npy = pat.npy_;
    pf.one = pat.one_;
    pf.bol = pat.bol_;
    pf.lbk = pat.lbk_;
    pf.lbm = pat.lbm_;
    pf.cut = pat.cut_;
    memcpy(pf.chr, pat.chr_, sizeof(pf.chr));
    memcpy(pf.bit, pat.bit_, sizeof(pf.bit));
    memcpy(pf.tap, pat.tap_, sizeof(pf.tap));
    memcpy(pf.pma, pat.pma_, sizeof(pf.pma));
    memcpy(pf.pmh, pat.pmh_, sizeof(pf.pmh));
    if (pat.bmd_ > 0)
      memcpy(pf.bms, pat.bms_, sizeof(pf.bms));
    for (int c = 0; c < 256; ++c)
    {
      if (pat.cbk_.test(c))
        pf.cbk[c >> 3] |= static_cast<uint8_t>(1 << (c & 7));
      if (pat.fst_.test(c))
        pf.fst[c >> 3] |= static_cast<uint8_t>(1 << (c & 7));
    }
    const uint32_t flags = (opt_.N ? UGX_OPT_N : 0u) | (opt_.W ? UGX_OPT_W : 0u);
    const char *dev = getenv("UGREP_B200_DEVICE");
    device_ = dev != NULL ? atoi(dev) : 0;
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const int rc = ugx_pattern_create(pat.opc_, pat.nop_, &pf, flags, device_, &shared_->pattern);
    if (rc == UGX_E_UNSUPPORTED)
      return unsupported(ugx_last_error());
    if (rc != UGX_OK)
      throw std::runtime_error(std::string("ugrep-b200: ") + ugx_last_error());
    const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    new_scanner();
    if (env_set("UGREP_B200_VERBOSE"))
      fprintf(stderr, "ugrep-b200: pattern served by libugrep_b200 (device %d; CUDA context + pattern upload %.3f s, scanner %.3f s)\n",
              device_, std::chrono::duration<double>(t1 - t0).count(),
              std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
  }

  void unsupported(const char *why)
  {
    if (env_set("UGREP_B200_REQUIRE"))
      throw std::runtime_error(std::string("ugrep-b200: pattern outside the GPU path's scope: ") + why);
    if (env_set("UGREP_B200_VERBOSE"))
      fprintf(stderr, "ugrep-b200: pattern stays with reflex::Matcher: %s\n", why);
  }

  void new_scanner()
  {
    if (ugx_scanner_create(device_, NULL, &scanner_) != UGX_OK)
      throw std::runtime_error(std::string("ugrep-b200: ") + ugx_last_error());
  }

  // the first find() on an input: everything in memory, one device scan, records stay on the device
  void scan_input()
  {
    fresh_ = false;
    const bool verbose = env_set("UGREP_B200_VERBOSE");
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    // Read the rest of the input in full, the way peek_more() does block by block (absmatcher.h:1612-1631): grow()
    // shifts out what lies before the current line (calling the caller's handler, keeping lno_ / num_ right) or
    // enlarges the buffer.  (AbstractMatcher::buffer() is not used: it assumes nothing has been read yet, and
    // Grep::init_is_binary has usually peeked at the first block by now, src/ugrep.cpp:3998-4017.)
    txt_ = buf_ + cur_;
    len_ = 0;
    while (!eof_)
    {
      if (end_ + blk_ + 1 >= max_)
        (void)grow();
      const size_t n = get(buf_ + end_, blk_ > 0 ? blk_ : max_ - end_ - 1);
      if (n == 0)
        eof_ = !wrap();
      else
        end_ += n;
    }
    // records number lines from the start of buf_; lno_ is the line of lpb_ (absmatcher.h:695-736)
    size_t before = 0;
    for (const char *s = buf_; s < lpb_; ++s)
      before += *s == '\n';
    const uint64_t base_line = static_cast<uint64_t>(lno_) - 1 - before;
    lpb_line_ = lno_;
    const ugx_match *dev = NULL;
    uint64_t n = 0;
    ugx_totals tot;
    memset(&tot, 0, sizeof(tot));
    const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    const int rc = ugx_find_all_device(scanner_, shared_->pattern, buf_, end_, 0, base_line, &dev, &n, &tot);
    if (rc != UGX_OK)
      throw std::runtime_error(std::string("ugrep-b200: ") + ugx_last_error());
    if (verbose)
    {
      const std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now();
      fprintf(stderr, "ugrep-b200: %zu bytes in memory after %.3f s; device scan %.3f s (kernels %.3f ms, %u launches), %llu records\n",
              static_cast<size_t>(end_), std::chrono::duration<double>(t1 - t0).count(),
              std::chrono::duration<double>(t2 - t1).count(), tot.kernel_ms, tot.launches, static_cast<unsigned long long>(n));
    }
    count_ = n;
    next_ = 0;
    batch_first_ = 0;
    batch_.clear();
    if (got_ == Const::BOB)
      got_ = Const::UNK; // this input has been seen (at_bob() is for the callers of SCAN)
  }

  const ugx_match& record(uint64_t i)
  {
    if (i < batch_first_ || i >= batch_first_ + batch_.size())
    {
      const uint64_t want = count_ - i < BATCH ? count_ - i : BATCH;
      batch_.resize(static_cast<size_t>(want));
      if (ugx_scanner_fetch(scanner_, batch_.data(), i, want) != UGX_OK)
        throw std::runtime_error(std::string("ugrep-b200: ") + ugx_last_error());
      batch_first_ = i;
    }
    return batch_[static_cast<size_t>(i - batch_first_)];
  }

  static const uint64_t BATCH = 1 << 16;
  std::shared_ptr<Shared> shared_;
  ugx_scanner *scanner_ = NULL;
  int device_ = 0;
  bool fresh_ = true;
  uint64_t count_ = 0, next_ = 0, batch_first_ = 0;
  uint64_t lpb_line_ = 1;
  std::vector<ugx_match> batch_;
};

#endif
