// efanna2e::Neighbor / NeighborPriorityQueue — host-side mirror of the candidate-list types the reference's CLIs include
// (include/efanna2e/neighbor.h:21-34 and :138-223 there), so that code written against the reference keeps compiling.
// The batched search does not use this class: its candidate list lives in shared memory inside the CUDA kernel
// (roargraph_b200/csrc/rg_search.cu, "candidate queue"), which reproduces exactly the behaviour documented here:
//   * order = (distance ascending, id ascending);
//   * bounded: a candidate worse than the last entry of a full list is dropped;
//   * the binary search for the insertion point abandons the insert when a probed entry that is not greater than the
//     candidate carries the candidate's id (:161) — for a deterministic distance function this rejects every repeat;
//   * closest_unexpanded() marks the entry under the cursor expanded and moves the cursor to the next unexpanded one;
//     an insert in front of the cursor pulls the cursor back.
#pragma once
#include <cstddef>
#include <mutex>
#include <vector>

#ifndef LIKELY
#define LIKELY(x) __builtin_expect(x, 1)
#define UNLIKELY(x) __builtin_expect(x, 0)
#endif

namespace efanna2e {

struct Neighbor {
  unsigned id;
  float distance;
  bool flag;  // true once expanded

  Neighbor() = default;
  Neighbor(unsigned id_, float distance_, bool flag_) : id(id_), distance(distance_), flag(flag_) {}
  bool operator<(const Neighbor &o) const { return distance < o.distance || (distance == o.distance && id < o.id); }
  bool operator==(const Neighbor &o) const { return id == o.id; }
};

typedef std::lock_guard<std::mutex> LockGuard;

class NeighborPriorityQueue {
 public:
  NeighborPriorityQueue() = default;
  explicit NeighborPriorityQueue(size_t capacity) : cap_(capacity), slots_(capacity + 1) {}

  void insert(const Neighbor &nbr) {
    if (n_ == cap_ && n_ > 0 && slots_[n_ - 1] < nbr) return;
    // upper bound of nbr in slots_[0, n_), giving up on an equal id met on the way
    size_t first = 0, len = n_;
    while (len > 0) {
      const size_t probe = first + (len >> 1);  // == (lo + hi) / 2 of the half-open range [first, first + len)
      if (nbr < slots_[probe]) {
        len = probe - first;
      } else {
        if (UNLIKELY(slots_[probe].id == nbr.id)) return;
        len -= probe + 1 - first;
        first = probe + 1;
      }
    }
    if (first >= slots_.size()) return;
    for (size_t i = n_; i > first; --i) slots_[i] = slots_[i - 1];  // the slack slot takes the evicted tail
    slots_[first] = Neighbor(nbr.id, nbr.distance, false);
    if (n_ < cap_) ++n_;
    if (first < cursor_) cursor_ = first;
  }

  Neighbor closest_unexpanded() {
    const size_t at = cursor_;
    slots_[at].flag = true;
    while (cursor_ < n_ && slots_[cursor_].flag) ++cursor_;
    return slots_[at];
  }

  bool has_unexpanded_node() const { return cursor_ < n_; }
  size_t size() const { return n_; }
  size_t capacity() const { return cap_; }
  void reserve(size_t capacity) {
    if (capacity + 1 > slots_.size()) slots_.resize(capacity + 1);
    cap_ = capacity;
  }
  Neighbor &operator[](size_t i) { return slots_[i]; }
  Neighbor operator[](size_t i) const { return slots_[i]; }
  void clear() { n_ = cursor_ = 0; }
  std::vector<Neighbor> &get_data() { return slots_; }

 private:
  size_t n_ = 0, cap_ = 0, cursor_ = 0;
  std::vector<Neighbor> slots_;
};

}  // namespace efanna2e
