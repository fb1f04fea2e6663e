// efanna2e::Distance operator surface of the B200 build (host side).
//
// Mirrors the reference's include/efanna2e/distance.h:15-20 (enum Metric, virtual Distance::compare), :39-89
// (DistanceL2) and :108-225 (DistanceInnerProduct, which returns MINUS the dot product) so that code written against
// the reference keeps compiling.  The batched search never calls these on the host: the same arithmetic runs inside
// the CUDA kernels (roargraph_b200/csrc/rg_search.cu).  compare() here exists for callers that evaluate a handful of
// pairs on the host (entry-point selection, result checks) and reproduces the rounding order the reference's AVX-512
// code has when built with its release flags: 16 interleaved partial sums with separate multiply and
// add, an 8-wide and a 4-wide fused tail, then pairwise folds.
#pragma once
#include <cmath>
#include <cstddef>

namespace efanna2e {

enum Metric { L2 = 0, INNER_PRODUCT = 1, FAST_L2 = 2, PQ = 3, COSINE = 4 };

class Distance {
 public:
  virtual float compare(const float *a, const float *b, unsigned length) const = 0;
  virtual ~Distance() {}
};

namespace detail {
template <bool kSquaredDiff>
inline float lane_reduce(const float *a, const float *b, unsigned n) {
  float s[16] = {0};
  unsigned i = 0;
  for (; i + 16 <= n; i += 16)
    for (unsigned l = 0; l < 16; ++l) {
      const float x = kSquaredDiff ? a[i + l] - b[i + l] : a[i + l];
      const float y = kSquaredDiff ? x : b[i + l];
      volatile float prod = x * y;  // keep the multiply and the add separately rounded
      s[l] = s[l] + prod;
    }
  float t[8];
  for (unsigned l = 0; l < 8; ++l) t[l] = s[l + 8] + s[l];
  if (n - i >= 8) {
    for (unsigned l = 0; l < 8; ++l) {
      const float x = kSquaredDiff ? a[i + l] - b[i + l] : a[i + l];
      const float y = kSquaredDiff ? x : b[i + l];
      t[l] = std::fmaf(x, y, t[l]);
    }
    i += 8;
  }
  float u[4];
  for (unsigned l = 0; l < 4; ++l) u[l] = t[l + 4] + t[l];
  if (n - i >= 4) {
    for (unsigned l = 0; l < 4; ++l) {
      const float x = kSquaredDiff ? a[i + l] - b[i + l] : a[i + l];
      const float y = kSquaredDiff ? x : b[i + l];
      u[l] = std::fmaf(x, y, u[l]);
    }
    i += 4;
  }
  for (unsigned l = 0; i + l < n; ++l) {
    const float x = kSquaredDiff ? a[i + l] - b[i + l] : a[i + l];
    const float y = kSquaredDiff ? x : b[i + l];
    u[l] = std::fmaf(x, y, u[l]);
  }
  return (u[0] + u[1]) + (u[2] + u[3]);
}
}  // namespace detail

class DistanceL2 : public Distance {
 public:
  float compare(const float *a, const float *b, unsigned length) const override {
    return detail::lane_reduce<true>(a, b, length);
  }
};

class DistanceInnerProduct : public Distance {
 public:
  float compare(const float *a, const float *b, unsigned length) const override {
    return -detail::lane_reduce<false>(a, b, length);
  }
};

}  // namespace efanna2e
