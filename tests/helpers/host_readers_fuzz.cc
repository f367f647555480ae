// ASan / UBSan harness for the host-only readers of libb200xgb.so (csrc/json.h JsonReader + UbjReader, csrc/legacy_io.cc):
// every file given on the command line is parsed as it is (must succeed, and writing it back must be a fixed point) and then
// `iters` times after a random truncation / byte flips / a clobbered 8-byte field (must end in a document or an exception,
// never in a sanitizer report).  Built and run by tests/test_host_sanitizers.py with g++ -fsanitize=address,undefined.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <random>
#include "booster.h"
using namespace b200;
long long b200::g_kernel_launches = 0;
static std::string slurp(const char* p) { std::ifstream f(p, std::ios::binary); return std::string((std::istreambuf_iterator<char>(f)), {}); }
int main(int argc, char** argv) {
  std::mt19937 rng(7);
  long ok = 0, rejected = 0;
  const int iters = argc > 1 ? atoi(argv[1]) : 1000;
  for (int a = 2; a < argc; ++a) {
    const std::string raw = slurp(argv[a]);
    const bool legacy = looks_like_legacy_binary(raw.data(), raw.size()) || legacy_serialized_model_section(raw.data(), raw.size()).first;
    for (int it = 0; it <= iters; ++it) {
      std::string b = raw;
      if (it > 0) {
        const int kind = it % 3;
        if (kind == 0) b.resize(rng() % (b.size() + 1));
        else if (kind == 1) for (int k = 0; k < 1 + (int)(rng() % 8); ++k) b[rng() % b.size()] = (char)rng();
        else { size_t off = rng() % (b.size() - 8); uint64_t v = rng(); v = (v << 32) | rng(); if (rng() & 1) v &= 0xffff; memcpy(&b[off], &v, 8); }
      }
      try {
        JPtr doc;
        if (legacy) { auto s = legacy_serialized_model_section(b.data(), b.size()); doc = s.first ? legacy_binary_to_doc(s.first, s.second) : legacy_binary_to_doc(b.data(), b.size()); }
        else if (b.size() > 1 && b[0] == '{' && (b[1] == '"' || b[1] == ' ' || b[1] == '\n')) doc = JsonReader(b.data(), b.size()).parse();
        else doc = UbjReader(reinterpret_cast<const unsigned char*>(b.data()), b.size()).parse();
        std::string o1, o2; ubj_write(*doc, &o1); json_write(*doc, &o2);
        if (it == 0) {                              // the pristine file: write -> read -> write is a fixed point in both encodings
          std::string r1, r2;
          ubj_write(*UbjReader(reinterpret_cast<const unsigned char*>(o1.data()), o1.size()).parse(), &r1);
          json_write(*JsonReader(o2.data(), o2.size()).parse(), &r2);
          if (r1 != o1 || r2 != o2) { fprintf(stderr, "round trip of %s is not a fixed point\n", argv[a]); return 2; }
        }
        ++ok;
      } catch (const std::exception& e) { if (it == 0) { fprintf(stderr, "pristine %s rejected: %s\n", argv[a], e.what()); return 3; } ++rejected; }
    }
  }
  printf("parsed %ld, rejected %ld\n", ok, rejected);
  return 0;
}
