// tests/cpp/facade_test.cpp — drives the C++ mirror of reflex::Matcher (include/ugrep_b200/matcher.hpp) with the
// caller loops of Grep::search (/root/reference/src/ugrep.cpp:10536-10586, :10857-11047) and prints what
// `ugrep -c`, `ugrep -c -o` and `ugrep -n -b -o` print.  Usage: facade_test PATTERN.ugxp MODE FILE   MODE: cl|cm|list|cl_loop|all
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "ugrep_b200/matcher.hpp"

int main(int argc, char** argv)
{
  if (argc != 4)
  {
    fprintf(stderr, "usage: facade_test PATTERN.ugxp cl|cm|list|cl_loop FILE\n");
    return 2;
  }
  try
  {
    ugx::Pattern pattern(argv[1]);
    ugx::Matcher matcher(pattern);
    std::ifstream f(argv[3], std::ios::binary);
    std::vector<char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const size_t nbytes = data.size();
    data.push_back('\0'); // the slot buffer() expects after the text
    matcher.buffer(data.data(), nbytes + 1);
    const std::string mode = argv[2];
    if (mode == "cl")
      printf("%zu\n", matcher.count_lines());
    else if (mode == "cm")
      printf("%zu\n", matcher.count_matches());
    else if (mode == "cl_loop")
    {
      // the reference's own -c loop: find, count, skip to the next line (src/ugrep.cpp:10567-10586)
      size_t lines = 0;
      while (matcher.find())
      {
        ++lines;
        matcher.skip_line();
      }
      printf("%zu\n", lines);
    }
    else if (mode == "all")
    {
      // every loop in one process: counts first, then the -n -b -o listing
      printf("cl=%zu\n", matcher.count_lines());
      printf("cm=%zu\n", matcher.count_matches());
      size_t lines = 0;
      while (matcher.find())
      {
        ++lines;
        matcher.skip_line();
      }
      printf("loop=%zu\n", lines);
      matcher.reset();
      while (matcher.find())
      {
        printf("%zu:%zu:", matcher.lineno(), matcher.first());
        fwrite(matcher.begin(), 1, matcher.size(), stdout);
        fputc('\n', stdout);
      }
    }
    else if (mode == "list")
    {
      while (matcher.find())
      {
        printf("%zu:%zu:", matcher.lineno(), matcher.first());
        fwrite(matcher.begin(), 1, matcher.size(), stdout);
        fputc('\n', stdout);
      }
    }
    else
      return 2;
    return 0;
  }
  catch (const ugx::regex_error& e)
  {
    fprintf(stderr, "%s (code %d)\n", e.what(), e.code());
    return e.code() == UGX_E_UNSUPPORTED ? 3 : 4;
  }
}
