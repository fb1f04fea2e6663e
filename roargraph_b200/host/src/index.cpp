// efanna2e::Index base: picks the Distance operator by metric like the reference (src/index.cpp:8-26 there):
// L2 -> DistanceL2, INNER_PRODUCT and COSINE -> DistanceInnerProduct (COSINE = IP on normalised rows).
#include "efanna2e/index.h"

namespace efanna2e {

Index::Index(const size_t dimension, const size_t n, Metric metric) : dimension_(dimension), nd_(n), metric_(metric) {
  if (metric == L2) distance_ = new DistanceL2();
  else distance_ = new DistanceInnerProduct();
}

Index::~Index() { delete distance_; }

}  // namespace efanna2e
