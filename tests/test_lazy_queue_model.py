"""Model check (CPU) of the search kernel's LAZY candidate queue (roargraph_b200/csrc/rg_search.cu: pending_append,
queue_flush, queue_pop_lazy — used for L_pq >= 256) against the reference's search loop over its eager
NeighborPriorityQueue (src/index_bipartite.cpp:2311-2420 with include/efanna2e/neighbor.h:138-223).

Both searches are restated in Python over the same random graph and distances and must expand the SAME nodes in the SAME
order and end with the same L best (distance, id) entries — the claim the kernel's comment block makes:
  * accepting against the sorted arrays' last entry is a superset of what the eager queue accepts,
  * the next node is the better of the cursor entry and the best unexpanded pending entry, and the search stops when that
    entry's exact rank (array entries before it + pending entries before it) is >= L,
  * an entry expanded while pending keeps its flag through the flush,
  * the entry point (not marked visited) evaluated a second time is dropped.
Distances are drawn from a handful of values so that ties broken by id are everywhere; L runs from tiny (constant overflow,
the rank check decides most pops) to larger than the graph; the pending block holds 31 entries as in the kernel.
The GPU tests check the CUDA implementation (tests/test_gpu_search.py::test_lazy_queue_equals_eager_queue)."""
import random

from hypothesis import given, settings, strategies as st

from test_queue_merge_model import SeqQueue, less

PEND_MAX = 31
EVENTS = {"flush_cut": 0, "pop_pending": 0, "rank_stop": 0, "ep_again": 0, "flush_mid_group": 0, "flag_kept": 0}


def make_graph(rng, n, deg, n_dist):
    dist = [rng.randrange(n_dist) * 0.5 if rng.random() < 0.7 else rng.random() * n_dist * 0.5 for _ in range(n)]
    adj = [rng.sample(range(n), min(n, rng.randint(1, deg))) for _ in range(n)]
    return dist, adj


def search_eager(dist, adj, ep, L):
    """:2311-2420: the entry point is inserted but NOT marked visited; every other id is evaluated once."""
    q = SeqQueue(L)
    visited = set()
    q.insert(dist[ep], ep)
    order = []
    while q.cur < q.size:
        u = q.data[q.cur][1]
        q.pop()
        order.append(u)
        for v in adj[u]:
            if v in visited:
                continue
            visited.add(v)
            q.insert(dist[v], v)
    return order, [(d, i) for d, i, _ in q.data[:q.size]]


class LazyQueue:
    def __init__(self, L):
        self.L = L
        self.main = []      # sorted [d, id, expanded]
        self.cur = 0
        self.pend = []      # unsorted [d, id, expanded]
        self.npu = 0        # unexpanded pending entries

    def flush(self):
        if not self.pend:
            return
        L = self.L
        block = sorted(self.pend, key=lambda e: (e[0], e[1]))
        size = len(self.main)

        def pend_before(e):
            return sum(1 for p in block if less((p[0], p[1]), (e[0], e[1])))

        cur_new = self.cur + pend_before(self.main[self.cur]) if self.cur < size else None
        out = {}
        for j, e in enumerate(self.main):
            nj = j + pend_before(e)
            if nj < L:
                out[nj] = e
        first = None
        for r, p in enumerate(block):
            ub = sum(1 for e in self.main if not less((p[0], p[1]), (e[0], e[1])))
            fin = ub + r
            if fin < L:
                out[fin] = p
                if not p[2] and (first is None or fin < first):
                    first = fin
        newsize = min(size + len(block), L)
        EVENTS["flush_cut"] += size + len(block) > L
        EVENTS["flag_kept"] += any(p[2] for p in block)
        assert sorted(out) == list(range(newsize))          # every slot below the new size has exactly one writer
        self.main = [out[j] for j in range(newsize)]
        cands = [c for c in (cur_new, first) if c is not None]
        self.cur = min(cands + [newsize])
        self.pend = []
        self.npu = 0

    def append_group(self, group, ep):
        """group: up to 8 (d, id) whose `pass` flag was computed against main's last entry BEFORE the call (evaluate())"""
        L = self.L
        EVENTS["ep_again"] += any(c[1] == ep for c in group)
        group = [c for c in group if c[1] != ep]            # drop_ep
        if not group:
            return
        if len(self.pend) + len(group) > PEND_MAX:
            EVENTS["flush_mid_group"] += 1
            self.flush()
            if len(self.main) == L:
                group = [c for c in group if less(c, (self.main[L - 1][0], self.main[L - 1][1]))]
            if not group:
                return
        for d, i in group:
            self.pend.append([d, i, False])
        self.npu += len(group)

    def pop(self):
        L = self.L
        has_a = self.cur < len(self.main)
        unexp = [p for p in self.pend if not p[2]]
        has_b = self.npu != 0
        assert has_b == bool(unexp)
        if not has_a and not has_b:
            return None
        a = (self.main[self.cur][0], self.main[self.cur][1]) if has_a else None
        b = min(((p[0], p[1]) for p in unexp), default=None)
        use_b = has_b and (not has_a or less(b, a))
        u = b if use_b else a
        if len(self.main) + len(self.pend) > L:
            cntp = sum(1 for p in self.pend if less((p[0], p[1]), u))
            posm = sum(1 for e in self.main if not less(u, (e[0], e[1]))) if use_b else self.cur
            if posm + cntp >= L:
                EVENTS["rank_stop"] += 1
                return None
        if use_b:
            EVENTS["pop_pending"] += 1
            for p in self.pend:
                if not p[2] and p[1] == u[1]:
                    p[2] = True
            self.npu -= 1
        else:
            self.main[self.cur][2] = True
            c = self.cur + 1
            while c < len(self.main) and self.main[c][2]:
                c += 1
            self.cur = c
        return u[1]


def search_lazy(dist, adj, ep, L):
    q = LazyQueue(L)
    visited = set()
    q.main = [[dist[ep], ep, False]]                        # the entry point goes straight into the (empty) sorted arrays
    order = []
    while q.cur < len(q.main) or q.npu:
        u = q.pop()
        if u is None:
            break
        order.append(u)
        fresh = []
        for v in adj[u]:
            if v in visited:
                continue
            visited.add(v)
            fresh.append((dist[v], v))
        for g0 in range(0, len(fresh), 8):                  # evaluate(): row groups of 8, `pass` against main's last entry
            group = fresh[g0:g0 + 8]
            if len(q.main) == L:
                group = [c for c in group if less(c, (q.main[L - 1][0], q.main[L - 1][1]))]
            q.append_group(group, ep)
    q.flush()
    return order, [(d, i) for d, i, _ in q.main]


@settings(max_examples=400, deadline=None)
@given(st.integers(0, 2**32 - 1), st.integers(2, 400), st.integers(1, 40), st.integers(1, 120), st.integers(1, 12))
def test_lazy_queue_search_equals_eager_queue_search(seed, n, deg, L, n_dist):
    rng = random.Random(seed)
    dist, adj = make_graph(rng, n, deg, n_dist)
    ep = rng.randrange(n)
    want_order, want_top = search_eager(dist, adj, ep, L)
    got_order, got_top = search_lazy(dist, adj, ep, L)
    assert got_order == want_order, (seed, n, deg, L)
    assert got_top == want_top, (seed, n, deg, L)


def test_the_model_runs_reach_every_branch():
    """a fixed sweep must exercise what the argument is about: flushes that cut at L, pops out of the pending block, stops
    decided by the exact rank, the entry point seen again, a flush forced in the middle of a hop, expanded flags carried over"""
    for k in EVENTS:
        EVENTS[k] = 0
    for seed in range(300):
        rng = random.Random(seed)
        n, deg, L = rng.choice([30, 150, 400]), rng.choice([3, 12, 40]), rng.choice([1, 3, 10, 40, 100])
        dist, adj = make_graph(rng, n, deg, rng.choice([2, 6, 12]))
        ep = rng.randrange(n)
        assert search_lazy(dist, adj, ep, L) == search_eager(dist, adj, ep, L), seed
    assert all(v > 0 for v in EVENTS.values()), EVENTS
