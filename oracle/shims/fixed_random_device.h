// TEST INFRASTRUCTURE ONLY — pre-included (g++ -include) in front of the reference's UNMODIFIED translation units when
// oracle/Makefile builds oracle/_ref/libref_roargraph_detrand.so.
//
// IndexBipartite::SearchBipartiteGraph (src/index_bipartite.cpp:282-356 of the reference) draws its 10 entry points from
//     std::random_device rd;  std::mt19937 gen(rd());  std::uniform_int_distribution<uint32_t> dis(0, u32_nd_ - 1);
// (:287-294), which cannot be seeded from outside.  This header swaps the ONE non-deterministic piece — the word that
// std::random_device returns — for a value the test harness sets per thread before each call; mt19937 and
// uniform_int_distribution stay libstdc++'s own, so the 10 ids are exactly what the reference computes from that seed word.
// Every other line of the reference is compiled as it lies under /root/reference.
#pragma once
#include <cstdint>
#include <random>  // the real std::random_device is declared before the macro below exists
#include <string>

namespace rg_detrand {
inline thread_local uint32_t next_seed = 0;  // returned by the next random_device::operator() of this thread
inline thread_local uint64_t calls = 0;      // how many words this thread has drawn (the harness checks it is 1 per query)
}  // namespace rg_detrand

namespace std {
class rg_fixed_random_device {
   public:
    using result_type = unsigned int;
    rg_fixed_random_device() {}
    explicit rg_fixed_random_device(const std::string &) {}
    result_type operator()() {
        ++rg_detrand::calls;
        return rg_detrand::next_seed;
    }
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xffffffffu; }
    double entropy() const noexcept { return 0.0; }
};
}  // namespace std

#define RG_DETRAND 1
#define random_device rg_fixed_random_device
