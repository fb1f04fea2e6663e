// Normalises seeded random rows with efanna2e::normalize<float> and prints the bits.  tests/test_host_types.py compiles
// it against the reference's include/efanna2e/util.h with the reference's release flags and against the host mirror's
// header with the mirror's flags, and requires identical output.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "efanna2e/util.h"

int main(int argc, char **argv) {
  const unsigned dim = argc > 1 ? (unsigned)atoi(argv[1]) : 64;
  const unsigned rows = argc > 2 ? (unsigned)atoi(argv[2]) : 100;
  uint64_t s = argc > 3 ? strtoull(argv[3], nullptr, 10) : 1;
  std::vector<float> x((size_t)rows * dim);
  for (auto &v : x) {  // splitmix64 -> roughly uniform in (-8, 8), a different scale per element
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    v = ((float)(z >> 40) / 16777216.0f - 0.5f) * 16.0f;
  }
  for (unsigned r = 0; r < rows; ++r) efanna2e::normalize<float>(x.data() + (size_t)r * dim, dim);
  for (size_t i = 0; i < x.size(); ++i) {
    uint32_t b;
    memcpy(&b, &x[i], 4);
    printf("%08x%c", b, (i + 1) % 16 ? ' ' : '\n');
  }
  printf("\n");
  return 0;
}
