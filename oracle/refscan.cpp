// oracle/refscan.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A small harness over the UNMODIFIED reference library (libreflex.a built by
// oracle/Makefile from /root/reference).  It does three things:
//
//   refscan dump  [popts] -o OUT.ugxp     compile a pattern exactly the way
//        ugrep hands it to reflex::Pattern (src/ugrep.cpp:8186-8604, :8849)
//        and write the compiled form (opcode words + prefilter fields,
//        include/reflex/pattern.h:1288-1334) in the UGXP container that
//        include/ugrep_b200.h documents.
//   refscan scan  MODE [popts] FILE       run the reference matcher in-place
//        (AbstractMatcher::buffer, absmatcher.h:542-591) with the caller loops
//        of Grep::search (src/ugrep.cpp:10536-10586, :10857-11047) and print
//        what `ugrep -c`, `ugrep -c -o`, `ugrep -n -b -o` print.
//   refscan isutf8 FILE...                 reflex::isutf8 (ugrep's binary-file test) per file
//   refscan bench MODE [popts] -J N -r R FILE   time R repetitions of the scan
//        over N line-aligned shards on N threads (one cloned matcher each, as
//        GrepWorker does, src/ugrep.cpp:4204-4215); prints seconds (best).
//
// popts:  -F  -i  -w  -U  -G  -Y  -e PATTERN (repeatable)  -f FILE
// MODE:   cl = count matching lines, cm = count matches, list = n:b:text
//
// Built with -fno-access-control so the dump can read Pattern's tables.
#include <reflex/matcher.h>
#include <reflex/pattern.h>
#include <reflex/simd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "../include/ugrep_b200.h"

struct POpts {
  bool F = false, i = false, w = false, U = false, G = false, Y = false;
  std::vector<std::string> pats;
  std::string file;
};

// CNF::quote, src/cnf.hpp:147-165
static void quote(std::string& p)
{
  if (p.empty())
    return;
  size_t from = 0, to;
  while ((to = p.find("\\E", from)) != std::string::npos)
  {
    p.insert(to + 2, "\\\\E\\Q");
    from = to + 7;
  }
  p.insert(0, "\\Q").append("\\E");
}

// the regex string as assembled by ugrep() for -e/-f/-F/-i (src/ugrep.cpp:8186-8362, 8585-8604)
static std::string assemble(const POpts& o)
{
  std::string regex;
  const char *bar = o.G ? "\\|" : "|";
  // patterns are split at newlines (CNF::split), each is an ALT term (CNF::adjoin)
  for (size_t n = 0; n < o.pats.size(); ++n)
  {
    std::string p = o.pats[n];
    size_t from = 0;
    while (true)
    {
      size_t nl = p.find('\n', from);
      std::string term = p.substr(from, nl == std::string::npos ? std::string::npos : nl - from);
      if (!term.empty())
      {
        if (o.F)
          quote(term);
        regex.append(term).append(bar);
      }
      if (nl == std::string::npos)
        break;
      from = nl + 1;
    }
  }
  if (!regex.empty())
  {
    regex.pop_back();
    if (o.G)
      regex.pop_back();
  }
  if (!o.file.empty())
  {
    bool fixed = o.F;
    if (!regex.empty())
    {
      fixed = false;
      regex.append(bar);
    }
    std::ifstream in(o.file.c_str());
    if (!in)
    {
      fprintf(stderr, "refscan: cannot read %s\n", o.file.c_str());
      exit(2);
    }
    std::string line;
    while (std::getline(in, line))
    {
      if (!line.empty() && line.back() == '\r')
        line.pop_back();
      if (!line.empty())
      {
        if (fixed)
          quote(line);
        regex.append(line).append(bar);
      }
    }
    if (!regex.empty())
    {
      regex.pop_back();
      if (o.G)
        regex.pop_back();
    }
  }
  std::string popt("(?m");
  if (o.i)
    popt.push_back('i');
  popt.push_back(')');
  regex.insert(0, popt);
  return regex;
}

static reflex::convert_flag_type cflags(const POpts& o)
{
  reflex::convert_flag_type f = reflex::convert_flag::notnewline;
  if (!o.U)
    f |= reflex::convert_flag::unicode;
  if (o.G)
    f |= reflex::convert_flag::basic;
  return f;
}

static std::string mopts(const POpts& o)
{
  std::string m;
  if (o.Y)
    m.push_back('N');
  if (o.w)
    m.push_back('W');
  return m;
}

static int parse_popts(int argc, char **argv, int i, POpts& o, std::vector<std::string>& rest)
{
  for (; i < argc; ++i)
  {
    std::string a = argv[i];
    if (a == "-F") o.F = true;
    else if (a == "-i") o.i = true;
    else if (a == "-w") o.w = true;
    else if (a == "-U") o.U = true;
    else if (a == "-G") o.G = true;
    else if (a == "-Y") o.Y = true;
    else if (a == "-e" && i + 1 < argc) o.pats.push_back(argv[++i]);
    else if (a == "-f" && i + 1 < argc) o.file = argv[++i];
    else rest.push_back(a);
  }
  // CNF::anchor (src/cnf.hpp:167-204): without -w/-x, a pattern that starts with ^ or ends with $ switches -Y on
  // (after -F quoting the pattern starts with \Q, so -F never triggers it; -f FILE lines are not checked)
  if (!o.F && !o.w)
    for (size_t n = 0; n < o.pats.size(); ++n)
      if (!o.pats[n].empty() && (o.pats[n][0] == '^' || o.pats[n][o.pats[n].size() - 1] == '$'))
        o.Y = true;
  return i;
}

static bool read_file(const std::string& name, std::vector<char>& data)
{
  FILE *f = fopen(name.c_str(), "rb");
  if (!f)
    return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  data.resize(static_cast<size_t>(n) + 1);
  size_t got = n > 0 ? fread(data.data(), 1, n, f) : 0;
  fclose(f);
  data[got] = '\0';
  data.resize(got + 1);
  return true;
}

static int do_dump(const POpts& o, const std::string& out)
{
  std::string regex = assemble(o);
  bool multiline = false;
  reflex::Pattern pattern(reflex::Matcher::convert(regex, cflags(o), &multiline), "r");
  ugx_prefilter pf;
  memset(&pf, 0, sizeof(pf));
  pf.len = static_cast<uint32_t>(pattern.len_);
  pf.min = static_cast<uint32_t>(pattern.min_);
  pf.pin = static_cast<uint32_t>(pattern.pin_);
  pf.lcp = pattern.lcp_;
  pf.lcs = pattern.lcs_;
  pf.bmd = static_cast<uint32_t>(pattern.bmd_);
  pf.npy = pattern.npy_;
  pf.one = pattern.one_;
  pf.bol = pattern.bol_;
  pf.lbk = pattern.lbk_;
  pf.lbm = pattern.lbm_;
  pf.cut = pattern.cut_;
  memcpy(pf.chr, pattern.chr_, 256);
  memcpy(pf.bit, pattern.bit_, 256);
  memcpy(pf.tap, pattern.tap_, sizeof(pf.tap));
  memcpy(pf.pma, pattern.pma_, sizeof(pf.pma));
  memcpy(pf.pmh, pattern.pmh_, sizeof(pf.pmh));
  if (pattern.bmd_ > 0) // bms_ is only written (and read) for the Boyer-Moore routines; left zero otherwise so that dumps are deterministic
    memcpy(pf.bms, pattern.bms_, 256);
  for (int c = 0; c < 256; ++c)
  {
    if (pattern.cbk_.test(c))
      pf.cbk[c >> 3] |= 1 << (c & 7);
    if (pattern.fst_.test(c))
      pf.fst[c >> 3] |= 1 << (c & 7);
  }
  ugx_file_header h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, UGX_FILE_MAGIC, 8);
  h.nop = pattern.nop_;
  h.regex_len = static_cast<uint32_t>(regex.size());
  h.prefilter_size = sizeof(pf);
  h.matcher_flags = (o.w ? UGX_OPT_W : 0) | (o.Y ? UGX_OPT_N : 0);
  FILE *f = fopen(out.c_str(), "wb");
  if (!f)
  {
    fprintf(stderr, "refscan: cannot write %s\n", out.c_str());
    return 2;
  }
  fwrite(&h, sizeof(h), 1, f);
  fwrite(&pf, sizeof(pf), 1, f);
  fwrite(pattern.opc_, sizeof(uint32_t), pattern.nop_, f);
  fwrite(regex.data(), 1, regex.size(), f);
  fclose(f);
  fprintf(stderr, "refscan: nop=%u len=%u min=%u pin=%u lcp=%u lcs=%u bmd=%u npy=%u one=%u bol=%u lbk=%u lbm=%u cut=%u\n",
      h.nop, pf.len, pf.min, pf.pin, pf.lcp, pf.lcs, pf.bmd, pf.npy, pf.one, pf.bol, pf.lbk, pf.lbm, pf.cut);
  return 0;
}

struct Result {
  size_t count = 0;
  std::string text;
};

// the three caller loops of Grep::search that the configs use
static void scan(reflex::Matcher& m, char *base, size_t nbytes, const std::string& mode, size_t line0, size_t off0, Result& r, bool emit)
{
  m.buffer(base, nbytes + 1); // src/ugrep.cpp:3939
  if (mode == "cl")
  {
    m.lineno_skip(true);
    while (m.find())
    {
      ++r.count;
      if (!m.at_bol())
        m.skip('\n');
    }
  }
  else if (mode == "cm")
  {
    m.lineno_skip(true);
    while (m.find())
      ++r.count;
  }
  else
  {
    char tmp[64];
    while (m.find())
    {
      ++r.count;
      if (emit)
      {
        int n = snprintf(tmp, sizeof(tmp), "%zu:%zu:", m.lineno() + line0, m.first() + off0);
        r.text.append(tmp, n);
        r.text.append(m.begin(), m.size());
        r.text.push_back('\n');
      }
      else
      {
        (void)m.lineno();
      }
    }
  }
}

int main(int argc, char **argv)
{
  if (argc < 2)
  {
    fprintf(stderr, "usage: refscan dump|scan|bench ...\n");
    return 2;
  }
  std::string cmd = argv[1];
  try
  {
    if (cmd == "dump")
    {
      POpts o;
      std::vector<std::string> rest;
      parse_popts(argc, argv, 2, o, rest);
      std::string out;
      for (size_t i = 0; i + 1 < rest.size(); ++i)
        if (rest[i] == "-o")
          out = rest[i + 1];
      if (out.empty())
      {
        fprintf(stderr, "refscan dump: -o OUT required\n");
        return 2;
      }
      return do_dump(o, out);
    }
    if (cmd == "isutf8")
    {
      // reflex::isutf8 (lib/simd.cpp:169) over every file named: one 0/1 per line
      for (int i = 2; i < argc; ++i)
      {
        std::vector<char> data;
        if (!read_file(argv[i], data))
          return 2;
        printf("%d\n", reflex::isutf8(data.data(), data.data() + data.size() - 1) ? 1 : 0);
      }
      return 0;
    }
    if (cmd == "scan" || cmd == "bench")
    {
      if (argc < 4)
        return 2;
      std::string mode = argv[2];
      POpts o;
      std::vector<std::string> rest;
      parse_popts(argc, argv, 3, o, rest);
      size_t jobs = 1, reps = 1;
      std::string file;
      for (size_t i = 0; i < rest.size(); ++i)
      {
        if (rest[i] == "-J" && i + 1 < rest.size())
          jobs = strtoul(rest[++i].c_str(), NULL, 10);
        else if (rest[i] == "-r" && i + 1 < rest.size())
          reps = strtoul(rest[++i].c_str(), NULL, 10);
        else
          file = rest[i];
      }
      std::vector<char> data;
      if (!read_file(file, data))
      {
        fprintf(stderr, "refscan: cannot read %s\n", file.c_str());
        return 2;
      }
      size_t nbytes = data.size() - 1;
      std::string regex = assemble(o);
      bool multiline = false;
      reflex::Pattern pattern(reflex::Matcher::convert(regex, cflags(o), &multiline), "r");
      std::string mo = mopts(o);
      if (cmd == "scan")
      {
        reflex::Matcher m(pattern, reflex::Input(), mo.c_str());
        Result r;
        scan(m, data.data(), nbytes, mode, 0, 0, r, true);
        if (mode == "list")
          fwrite(r.text.data(), 1, r.text.size(), stdout);
        else
          printf("%zu\n", r.count);
        return r.count > 0 ? 0 : 1;
      }
      // bench: N line-aligned shards, one thread + one matcher each
      if (jobs < 1)
        jobs = 1;
      std::vector<size_t> cut(jobs + 1, nbytes);
      cut[0] = 0;
      for (size_t j = 1; j < jobs; ++j)
      {
        size_t p = nbytes / jobs * j;
        if (p < cut[j - 1])
          p = cut[j - 1];
        const char *q = p < nbytes ? static_cast<const char*>(memchr(data.data() + p, '\n', nbytes - p)) : NULL;
        cut[j] = q ? static_cast<size_t>(q - data.data()) + 1 : nbytes;
      }
      // in-place mode needs a writable NUL slot after each shard: copy shards (outside the timed region)
      std::vector<std::vector<char> > shard(jobs);
      for (size_t j = 0; j < jobs; ++j)
      {
        shard[j].assign(data.begin() + cut[j], data.begin() + cut[j + 1]);
        shard[j].push_back('\0');
      }
      double best = 1e30;
      size_t total = 0;
      for (size_t rep = 0; rep < reps; ++rep)
      {
        std::vector<Result> res(jobs);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (size_t j = 0; j < jobs; ++j)
          th.emplace_back([&, j]() {
            reflex::Matcher m(pattern, reflex::Input(), mo.c_str());
            scan(m, shard[j].data(), shard[j].size() - 1, mode, 0, 0, res[j], false);
          });
        for (auto& t : th)
          t.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best)
          best = dt;
        total = 0;
        for (size_t j = 0; j < jobs; ++j)
          total += res[j].count;
      }
      printf("{\"seconds\": %.6f, \"bytes\": %zu, \"threads\": %zu, \"count\": %zu}\n", best, nbytes, jobs, total);
      return 0;
    }
  }
  catch (reflex::regex_error& e)
  {
    fprintf(stderr, "refscan: regex error: %s\n", e.what());
    return 2;
  }
  fprintf(stderr, "refscan: unknown command %s\n", cmd.c_str());
  return 2;
}
