"""Model check (CPU) of the kNN re-rank kernel's list update (roargraph_b200/csrc/rg_knn.cu: knn_merge_fast_kernel, merge by
rank): the K-entry list of a query is kept sorted, the entries kept from a chunk are sorted alone and every entry of either list
lands at its own index plus the number of entries of the OTHER list below it (count_below = one binary search); positions
>= K fall off, the writer of position K - 1 tightens tau.  Restated in Python over packed (ordered-distance << 32 | id) words,
it must give exactly the first K of a full sort of both lists — what the heap scan of compute_groundtruth.cpp:203-245 keeps —
including short lists (the all-ones "empty" word repeated), a kept block larger than K, and equal distances ordered by id.
The GPU tests check the CUDA implementation (tests/test_gpu_knn.py::test_knn_merge_by_rank_equals_full_sort); this checks
the argument it rests on: the words are unique, so ranks in the two lists add up to a permutation."""
import random

from hypothesis import given, settings, strategies as st

EMPTY = (1 << 64) - 1


def count_below(a, n, key):
    lo, hi = 0, n
    while lo < hi:
        mid = (lo + hi) >> 1
        if a[mid] < key:
            lo = mid + 1
        else:
            hi = mid
    return lo


def bitonic_sort(a):
    """the network of bitonic_sort_smem / bitonic_sort_warp (m a power of two)"""
    m = len(a)
    size = 2
    while size <= m:
        stride = size >> 1
        while stride > 0:
            for i in range(m >> 1):
                lo = 2 * i - (i & (stride - 1))
                hi = lo + stride
                up = (lo & size) == 0
                if (a[lo] > a[hi]) == up:
                    a[lo], a[hi] = a[hi], a[lo]
            stride >>= 1
        size <<= 1


def merge_by_rank(top, kept, K, tau):
    """-> (new list, new tau-word or None) exactly as the kernel's threads compute them, every thread independent"""
    m2 = 1
    while m2 < len(kept):
        m2 <<= 1
    sk = list(kept) + [EMPTY] * (m2 - len(kept))
    bitonic_sort(sk)
    st_ = list(top)
    out = list(top)       # global memory: positions nobody writes keep their old word (must not happen for pos < K)
    written = [0] * K
    new_tau = tau
    for i in range(len(kept) + K):
        fresh = i < len(kept)
        key = sk[i] if fresh else st_[i - len(kept)]
        pos = i + count_below(st_, K, key) if fresh else (i - len(kept)) + count_below(sk, len(kept), key)
        if pos < K:
            out[pos] = key
            written[pos] += 1
            if pos == K - 1 and key != EMPTY:
                new_tau = min(new_tau, key >> 32)
    return out, written, new_tau


def words(rng, n, n_dist):
    """n unique words with few distinct distances (many ties broken by id)"""
    seen = set()
    while len(seen) < n:
        seen.add((rng.randrange(n_dist) << 32) | rng.randrange(1 << 20))
    return list(seen)


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 2**32 - 1), st.integers(1, 40), st.integers(0, 90), st.integers(0, 40), st.integers(1, 50))
def test_merge_by_rank_is_the_full_sort(seed, K, n_kept, n_filled, n_dist):
    rng = random.Random(seed)
    n_filled = min(n_filled, K)
    pool = words(rng, n_filled + n_kept, n_dist)
    top = sorted(pool[:n_filled]) + [EMPTY] * (K - n_filled)
    kept = pool[n_filled:]
    rng.shuffle(kept)
    if not kept:
        return  # the kernel skips the update when nothing was kept
    tau0 = 1 << 40
    got, written, tau = merge_by_rank(top, kept, K, tau0)
    want = sorted(top + kept)[:K]
    assert got == want
    assert written == [1] * K                      # every position below K has exactly one writer
    assert tau == (min(tau0, want[K - 1] >> 32) if want[K - 1] != EMPTY else tau0)


def test_bitonic_network_sorts_with_padding():
    rng = random.Random(5)
    for m in (1, 2, 4, 8, 16, 32, 64, 128):
        for fill in (0, 1, m // 2, m):
            a = words(rng, fill, 7) + [EMPTY] * (m - fill)
            rng.shuffle(a)
            b = list(a)
            bitonic_sort(b)
            assert b == sorted(a)
