"""Model check (CPU) of the search kernel's merge insert (roargraph_b200/csrc/rg_search.cu: queue_insert_group) against
the reference's one-by-one NeighborPriorityQueue::insert (include/efanna2e/neighbor.h:150-183): the same algorithm the
kernel runs per row group, restated in Python over plain lists, must leave the queue — entries, expanded flags, size and
cursor — exactly as sequential inserts do, for groups of up to 8 candidates with distinct ids, including distance ties,
exact duplicates of queued entries, queues that overflow inside the group, and expanded entries around the cursor.
The GPU parity tests check the CUDA implementation; this checks the argument it rests on."""
import random

from hypothesis import given, settings, strategies as st


def less(a, b):  # neighbor.h:29-31 on (distance, id)
    return a[0] < b[0] or (a[0] == b[0] and a[1] < b[1])


class SeqQueue:
    """neighbor.h:138-223, literally."""

    def __init__(self, cap):
        self.cap, self.size, self.cur = cap, 0, 0
        self.data = [None] * (cap + 1)  # (dist, id, flag)

    def insert(self, d, i):
        if self.size == self.cap and less(self.data[self.size - 1], (d, i)):
            return
        lo, hi = 0, self.size
        while lo < hi:
            mid = (lo + hi) >> 1
            if less((d, i), self.data[mid]):
                hi = mid
            elif self.data[mid][1] == i:
                return
            else:
                lo = mid + 1
        if lo < self.cap:
            self.data[lo + 1:self.size + 1] = self.data[lo:self.size]
        self.data[lo] = (d, i, False)
        if self.size < self.cap:
            self.size += 1
        if lo < self.cur:
            self.cur = lo

    def pop(self):
        e = self.data[self.cur]
        self.data[self.cur] = (e[0], e[1], True)
        while self.cur < self.size and self.data[self.cur][2]:
            self.cur += 1

    def state(self):
        return self.data[:self.size], self.size, self.cur


class MergeQueue(SeqQueue):
    """queue_insert_group: pre-filter against the worst entry, per-candidate upper bound + duplicate test, ranks among the
    accepted, one shifting pass, placement, cursor = min(cursor, smallest landing slot)."""

    def insert_group(self, group):
        L = self.cap
        cands = [c for c in group if not (self.size == L and not less(c, self.data[L - 1]))]
        if len(cands) == 1:
            return self.insert(*cands[0])
        acc = []
        for d, i in cands:
            lo, hi = 0, self.size
            while lo < hi:
                mid = (lo + hi) >> 1
                if less((d, i), self.data[mid]):
                    hi = mid
                else:
                    lo = mid + 1
            if lo > 0 and self.data[lo - 1][1] == i and self.data[lo - 1][0] == d:
                continue
            acc.append((d, i, lo))
        if not acc:
            return
        old = self.data[:self.size]
        minpos = min(p for _, _, p in acc)
        newsize = min(self.size + len(acc), L)
        out = list(self.data)
        for j in range(self.size - 1, minpos - 1, -1):
            nj = j + sum(1 for _, _, p in acc if p <= j)
            if nj < L:
                out[nj] = old[j]
        for d, i, p in acc:
            fin = p + sum(1 for d2, i2, _ in acc if less((d2, i2), (d, i)))
            if fin < L:
                out[fin] = (d, i, False)
        self.data = out
        self.size = newsize
        self.cur = min(self.cur, minpos)


@settings(max_examples=400, deadline=None)
@given(st.integers(1, 24), st.integers(0, 2**32 - 1), st.integers(1, 60))
def test_merge_insert_equals_sequential_inserts(cap, seed, n_groups):
    rng = random.Random(seed)
    a, b = SeqQueue(cap), MergeQueue(cap)
    n_ids = rng.choice([6, 20, 200])
    dist_of = {i: rng.choice([0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0]) if rng.random() < 0.5 else rng.random() * 3 for i in range(n_ids)}
    for _ in range(n_groups):
        if rng.random() < 0.3 and a.cur < a.size:
            a.pop()
            b.pop()
        group = [(dist_of[i], i) for i in rng.sample(range(n_ids), rng.randint(1, min(8, n_ids)))]  # distinct ids
        for d, i in group:
            a.insert(d, i)
        b.insert_group(group)
        assert a.state() == b.state(), (cap, seed)
