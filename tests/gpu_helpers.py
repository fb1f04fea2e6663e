"""Helpers shared by the -m gpu tests (case construction with the oracle / reference as CHECKERS only)."""
import os
import tempfile

import numpy as np

from roargraph_b200 import formats as F, synth


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def assert_same_search(got, want, what=""):
    ids, dists, cmps, hops = got
    assert (ids == want["ids"]).all(), f"{what}: ids differ in {np.count_nonzero((ids != want['ids']).any(1))} queries"
    assert (bits(dists) == bits(want["dists"])).all(), f"{what}: distances not bit-equal"
    assert (cmps == want["cmps"]).all(), f"{what}: cmps differ"
    assert (hops == want["hops"]).all(), f"{what}: hops differ"


def random_graph(rng, nd, deg_lo, deg_hi, dup_frac=0.0, self_loops=False, zero_frac=0.0):
    """Random digraph as CSR; optionally with repeated ids inside a row, self loops and empty rows."""
    degs = rng.integers(deg_lo, deg_hi + 1, nd)
    if zero_frac > 0:
        degs[rng.random(nd) < zero_frac] = 0
    row_off = np.zeros(nd + 1, dtype=np.uint64)
    row_off[1:] = np.cumsum(degs)
    nbrs = rng.integers(0, nd, int(row_off[-1])).astype(np.uint32)
    if dup_frac > 0:
        for u in rng.choice(nd, max(1, int(nd * dup_frac)), replace=False):
            b, e = int(row_off[u]), int(row_off[u + 1])
            if e - b >= 2:
                nbrs[b + 1] = nbrs[b]
    if self_loops:
        for u in range(0, nd, 7):
            b, e = int(row_off[u]), int(row_off[u + 1])
            if e > b:
                nbrs[b] = u
    return row_off, nbrs


def bipartite_from_knn(knn_ids, nd, M_pjbp):
    """qbaseNNbipartite (src/index_bipartite.cpp:235-280): query node sq+nd -> its first M_pjbp NNs except the top-1;
    base node top-1 -> query node."""
    nsq = knn_ids.shape[0]
    rows = [[] for _ in range(nd + nsq)]
    for sq in range(nsq):
        nn = knn_ids[sq][:M_pjbp]
        tgt = int(nn[0])
        rows[nd + sq] = [int(x) for x in nn if int(x) != tgt]
        rows[tgt].append(nd + sq)
    row_off = np.zeros(nd + nsq + 1, dtype=np.uint64)
    row_off[1:] = np.cumsum([len(r) for r in rows])
    nbrs = np.array([x for r in rows for x in r], dtype=np.uint32)
    return row_off, nbrs


def build_reference_case(nb, ntrain, ntest, dim, metric, M_sq=100, M_pjbp=35, L_pjpq=500, seed=1234, threads=0):
    """Synthetic data -> kNN on the GPU (rg.knn) -> index built by the REFERENCE's BuildRoarGraph (oracle/_ref).
    Returns dict(base, queries (padded/normalised as the reference's search sees them), ep, row_off, nbrs, ref handle)."""
    import roargraph_b200 as rg
    from oracle.oracle import Reference, run_ref_build_roargraph
    R = Reference()
    threads = threads or R.num_procs()
    base, train, test = synth.make_numpy(nb, ntrain, ntest, dim, seed=seed)
    d = tempfile.mkdtemp(prefix="rgcase_")
    bf, gf, kf = os.path.join(d, "base.fbin"), os.path.join(d, "index.rg"), os.path.join(d, "train.gt")
    tf = os.path.join(d, "train.fbin")
    F.write_fbin(bf, base)
    F.write_fbin(tf, train)
    ids, dists = rg.knn(base, train, M_sq, metric)
    if metric == "ip":
        dists = -dists
    F.write_truthset(kf, ids.astype(np.uint32), dists)
    run_ref_build_roargraph(bf, tf, kf, gf, metric, M_sq, M_pjbp, L_pjpq, threads)
    ep, row_off, nbrs = F.read_projection_index(gf)
    pdim = (dim + 7) // 8 * 8
    h = R.index_load(bf, gf, metric, pdim, threads)
    q = F.pad_rows(test)
    if metric == "cosine":
        q = R.normalize_rows(q)
    b = R.index_base(h, nb, pdim).copy()
    return dict(R=R, h=h, base=b, queries=q, ep=ep, row_off=row_off, nbrs=nbrs, threads=threads, raw_base=base,
                raw_test=test, train_knn=ids, dir=d)
