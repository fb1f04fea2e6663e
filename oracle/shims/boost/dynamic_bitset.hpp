// Minimal stand-in for <boost/dynamic_bitset.hpp> (Boost is not installed in this image).
// TEST INFRASTRUCTURE ONLY: lets the unmodified reference translation units under
// /root/reference compile for oracle/_ref. Covers exactly the members the reference uses
// (ctor (n[,val]), reset/reserve/resize/test/set/operator[]/size, copy).
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
template <class Block = unsigned long>
class dynamic_bitset {
    std::vector<bool> bits_;
   public:
    dynamic_bitset() = default;
    explicit dynamic_bitset(std::size_t n, unsigned long val = 0) : bits_(n, false) {
        for (std::size_t i = 0; i < n && i < 8 * sizeof(unsigned long); ++i) bits_[i] = (val >> i) & 1ul;
    }
    void reset() { bits_.assign(bits_.size(), false); }
    void reserve(std::size_t n) { bits_.reserve(n); }
    void resize(std::size_t n, bool v = false) { bits_.resize(n, v); }
    bool test(std::size_t i) const { return bits_[i]; }
    dynamic_bitset &set(std::size_t i, bool v = true) { bits_[i] = v; return *this; }
    bool operator[](std::size_t i) const { return bits_[i]; }
    std::vector<bool>::reference operator[](std::size_t i) { return bits_[i]; }
    std::size_t size() const { return bits_.size(); }
};
}  // namespace boost
