// Drives efanna2e::NeighborPriorityQueue through a seeded random schedule of inserts / closest_unexpanded calls and prints
// a trace of everything observable.  tests/test_host_types.py compiles it twice — against the reference's
// include/efanna2e/neighbor.h and against the host mirror's — and requires identical traces.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "efanna2e/neighbor.h"

static uint64_t s_state;
static uint32_t next_u32() {  // splitmix64
  uint64_t z = (s_state += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

int main(int argc, char **argv) {
  const unsigned capacity = argc > 1 ? (unsigned)atoi(argv[1]) : 16;
  const unsigned n_ops = argc > 2 ? (unsigned)atoi(argv[2]) : 2000;
  const unsigned id_range = argc > 3 ? (unsigned)atoi(argv[3]) : 64;
  s_state = argc > 4 ? strtoull(argv[4], nullptr, 10) : 1;
  efanna2e::NeighborPriorityQueue q(capacity);
  for (unsigned op = 0; op < n_ops; ++op) {
    const uint32_t r = next_u32();
    if ((r & 7) == 0 && q.has_unexpanded_node()) {
      efanna2e::Neighbor n = q.closest_unexpanded();
      printf("pop %u %a\n", n.id, n.distance);
    } else {
      const unsigned id = next_u32() % id_range;
      // a deterministic distance per id (what a real distance function gives), coarse enough to tie between ids
      const float d = (float)((id * 2654435761u >> 20) % 37) * 0.25f;
      q.insert(efanna2e::Neighbor(id, d, false));
    }
    if (op % 97 == 0 || op + 1 == n_ops) {
      printf("state size=%zu cap=%zu more=%d :", q.size(), q.capacity(), (int)q.has_unexpanded_node());
      for (size_t i = 0; i < q.size(); ++i) printf(" %u/%a/%d", q[i].id, q[i].distance, (int)q[i].flag);
      printf("\n");
    }
  }
  q.clear();
  q.reserve(capacity * 2);
  printf("after clear: size=%zu cap=%zu more=%d\n", q.size(), q.capacity(), (int)q.has_unexpanded_node());
  return 0;
}
