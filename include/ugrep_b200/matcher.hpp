// ugrep_b200/matcher.hpp — host-side C++ mirror of the reference interface for the buffer-scan path,
// header-only over the C ABI (ugrep_b200.h).  Names, argument meaning and error behaviour follow
// reflex::Pattern / reflex::Matcher for the members ugrep's search loops use on this path
// (/root/reference/include/reflex/absmatcher.h, matcher.h, pattern.h; call sites src/ugrep.cpp:10536-11047):
//
//   reflex::Pattern(code, pred)              -> ugx::Pattern(opc, nop, prefilter, options)   (pattern.h:151-159)
//   reflex::regex_error                      -> ugx::regex_error                              (pattern.cpp:162-169)
//   Matcher(pattern, input, "N|W|...")       -> ugx::Matcher(pattern, options)                (absmatcher.h:354-388)
//   matcher.buffer(base, size)               -> same: size = nbytes + 1, caller keeps base alive (absmatcher.h:542-591)
//   matcher.find()                           -> accept index of the next match, 0 at the end  (absmatcher.h:1413)
//   begin() size() text() first() last()     -> span and byte offsets of the current match    (absmatcher.h:901-905)
//   lineno() lines()? bol()? at_bol()? ...   -> lineno() only: the state Output::header reads  (absmatcher.h:695-766)
//   skip('\n')                               -> skip_line(): what `ugrep -c` does after a hit (absmatcher.h:1198-1220)
//
// The scan itself runs on the GPU in ONE call per buffer (ugx_find_all_device); find() replays the device-
// produced records in input order, fetching them in batches.  count_lines() / count_matches() are the bulk
// forms of the `-c` / `-c -o` loops.  There is no CPU fallback: any CUDA failure throws.
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ugrep_b200.h"

namespace ugx {

class regex_error : public std::runtime_error {
 public:
  regex_error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; } // UGX_E_*
 private:
  int code_;
};

inline void check(int rc)
{
  if (rc != UGX_OK)
    throw regex_error(rc, std::string("ugrep_b200: ") + ugx_last_error());
}

inline uint32_t parse_options(const char* opt)
{
  uint32_t f = 0;
  for (; opt != nullptr && *opt != '\0'; ++opt)
  {
    if (*opt == 'N')
      f |= UGX_OPT_N;
    else if (*opt == 'W')
      f |= UGX_OPT_W;
  }
  return f;
}

// A compiled pattern: immutable after construction, shareable between threads (as reflex::Pattern).
class Pattern {
 public:
  Pattern(const uint32_t* opc, uint32_t nop, const ugx_prefilter& pf, const char* matcher_options = "", int device = 0)
  {
    check(ugx_pattern_create(opc, nop, &pf, parse_options(matcher_options), device, &p_));
  }
  explicit Pattern(const std::string& ugxp_file, int device = 0) { check(ugx_pattern_load(ugxp_file.c_str(), device, &p_)); }
  Pattern(const Pattern&) = delete;
  Pattern& operator=(const Pattern&) = delete;
  ~Pattern() { ugx_pattern_destroy(p_); }
  const ugx_pattern* handle() const { return p_; }
  ugx_pattern_info info() const
  {
    ugx_pattern_info i;
    check(ugx_pattern_info_get(p_, &i));
    return i;
  }
  size_t nodes() const { return info().states; }  // as Pattern::nodes()
  size_t words() const { return info().nop; }     // as Pattern::words()
 private:
  ugx_pattern* p_ = nullptr;
};

// One matcher per host thread (as reflex::Matcher: never shared, no locks inside).
class Matcher {
 public:
  static constexpr size_t BATCH = 1 << 16;

  explicit Matcher(const Pattern& pattern, int device = 0, void* stream = nullptr) : pat_(&pattern)
  {
    check(ugx_scanner_create(device, stream, &s_));
  }
  Matcher(const Matcher&) = delete;
  Matcher& operator=(const Matcher&) = delete;
  ~Matcher() { ugx_scanner_destroy(s_); }

  // clone(): same pattern, own scanner (GrepWorker::matcher_clone, src/ugrep.cpp:4204-4225)
  Matcher* clone(int device = 0, void* stream = nullptr) const { return new Matcher(*pat_, device, stream); }

  // in-place buffer: size counts the would-be NUL slot, i.e. nbytes + 1 (src/ugrep.cpp:3939)
  Matcher& buffer(char* base, size_t size)
  {
    if (base == nullptr || size == 0)
      throw regex_error(UGX_E_INVALID, "Matcher::buffer: null buffer");
    base_ = base;
    end_ = size - 1;
    reset();
    return *this;
  }
  void reset()
  {
    scanned_ = false;
    next_ = 0;
    nrec_ = 0;
    batch_.clear();
    batch_first_ = 0;
    cur_ = nullptr;
    skip_to_ = 0;
  }

  // the next match: its accept index (1-based alternative), or 0 when there is none
  size_t find()
  {
    if (!scanned_)
      scan();
    for (;;)
    {
      if (next_ >= nrec_)
      {
        cur_ = nullptr;
        return 0;
      }
      if (next_ < batch_first_ || next_ >= batch_first_ + batch_.size())
        fetch(next_);
      const ugx_match& m = batch_[next_ - batch_first_];
      ++next_;
      if (m.offset < skip_to_)
        continue; // skipped by skip_line()
      cur_ = &m;
      return m.cap;
    }
  }
  // after a hit: continue with the next line (what `ugrep -c` does: matcher->skip('\n'))
  void skip_line()
  {
    if (cur_ == nullptr)
      return;
    size_t p = static_cast<size_t>(cur_->offset) + cur_->len;
    while (p < end_ && base_[p] != '\n')
      ++p;
    skip_to_ = p + 1;
  }
  size_t accept() const { return cur_ ? cur_->cap : 0; }
  const char* begin() const { return cur_ ? base_ + cur_->offset : base_; }
  const char* text() const { return begin(); }
  size_t size() const { return cur_ ? cur_->len : 0; }
  size_t first() const { return cur_ ? static_cast<size_t>(cur_->offset) : 0; }
  size_t last() const { return first() + size(); }
  size_t lineno() const { return cur_ ? static_cast<size_t>(cur_->line) : 1; }
  std::string str() const { return std::string(begin(), size()); }

  // bulk forms of the counting loops
  size_t count_lines()
  {
    ugx_totals t;
    check(ugx_count_lines(s_, pat_->handle(), base_, end_, &t));
    last_ = t;
    return static_cast<size_t>(t.matches);
  }
  size_t count_matches()
  {
    ugx_totals t;
    check(ugx_count_matches(s_, pat_->handle(), base_, end_, &t));
    last_ = t;
    return static_cast<size_t>(t.matches);
  }
  const ugx_totals& totals() const { return last_; }
  size_t matches() const { return nrec_; }
  void set_option(const char* name, int value) { check(ugx_scanner_set_option(s_, name, value)); }

 private:
  void scan()
  {
    const ugx_match* dev = nullptr;
    uint64_t n = 0;
    check(ugx_find_all_device(s_, pat_->handle(), base_, end_, 0, 0, &dev, &n, &last_));
    nrec_ = static_cast<size_t>(n);
    scanned_ = true;
    next_ = 0;
  }
  void fetch(size_t first)
  {
    const size_t count = nrec_ - first < BATCH ? nrec_ - first : BATCH;
    batch_.resize(count);
    check(ugx_scanner_fetch(s_, batch_.data(), first, count));
    batch_first_ = first;
  }

  const Pattern* pat_;
  ugx_scanner* s_ = nullptr;
  char* base_ = nullptr;
  size_t end_ = 0;
  bool scanned_ = false;
  size_t next_ = 0, nrec_ = 0, batch_first_ = 0, skip_to_ = 0;
  std::vector<ugx_match> batch_;
  const ugx_match* cur_ = nullptr;
  ugx_totals last_{};
};

} // namespace ugx
