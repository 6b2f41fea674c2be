// tests/cpp/sharded_test.cpp — a plain C++ caller of ugx_sharded_* (include/ugrep_b200.h): reads a UGXP pattern file and
// a text file, shards the text over the named devices in ONE process and prints the totals of `ugrep -c`,
// `ugrep -c -o` and checksums of the `ugrep -o -n -b` records.  Usage: sharded_test PATTERN.ugxp FILE DEV[,DEV...]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "ugrep_b200.h"

int main(int argc, char** argv)
{
  if (argc != 4)
  {
    fprintf(stderr, "usage: sharded_test PATTERN.ugxp FILE DEV[,DEV...]\n");
    return 2;
  }
  std::ifstream pf(argv[1], std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(pf)), std::istreambuf_iterator<char>());
  ugx_file_header h;
  if (raw.size() < sizeof(h) + sizeof(ugx_prefilter))
    return 2;
  memcpy(&h, raw.data(), sizeof(h));
  if (memcmp(h.magic, UGX_FILE_MAGIC, 8) != 0 || h.prefilter_size != sizeof(ugx_prefilter))
    return 2;
  ugx_prefilter pre;
  memcpy(&pre, raw.data() + sizeof(h), sizeof(pre));
  std::vector<uint32_t> opc(h.nop);
  memcpy(opc.data(), raw.data() + sizeof(h) + sizeof(pre), 4ull * h.nop);
  std::ifstream tf(argv[2], std::ios::binary);
  std::vector<char> text((std::istreambuf_iterator<char>(tf)), std::istreambuf_iterator<char>());
  std::vector<int> devs;
  for (char* tok = strtok(argv[3], ","); tok != nullptr; tok = strtok(nullptr, ","))
    devs.push_back(atoi(tok));
  ugx_sharded* s = nullptr;
  if (ugx_sharded_create(opc.data(), h.nop, &pre, h.matcher_flags, devs.data(), static_cast<int>(devs.size()), &s) != UGX_OK)
  {
    fprintf(stderr, "create: %s\n", ugx_sharded_last_error());
    return 1;
  }
  ugx_totals lines, matches, recs;
  uint64_t n = 0;
  if (ugx_sharded_scan(s, text.data(), text.size(), UGX_MODE_LINES, nullptr, 0, nullptr, &lines, nullptr) != UGX_OK ||
      ugx_sharded_scan(s, text.data(), text.size(), UGX_MODE_MATCHES, nullptr, 0, nullptr, &matches, nullptr) != UGX_OK)
  {
    fprintf(stderr, "scan: %s\n", ugx_sharded_last_error());
    return 1;
  }
  std::vector<ugx_match> out(matches.matches + 1);
  if (ugx_sharded_scan(s, text.data(), text.size(), UGX_MODE_RECORDS, out.data(), out.size(), &n, &recs, nullptr) != UGX_OK)
  {
    fprintf(stderr, "records: %s\n", ugx_sharded_last_error());
    return 1;
  }
  unsigned long long osum = 0, lsum = 0;
  for (uint64_t i = 0; i < n; ++i)
  {
    osum += out[i].offset;
    lsum += out[i].line;
  }
  printf("lines=%llu matches=%llu records=%llu offset_sum=%llu line_sum=%llu\n", (unsigned long long)lines.matches,
         (unsigned long long)matches.matches, (unsigned long long)n, osum, lsum);
  ugx_sharded_destroy(s);
  return 0;
}
