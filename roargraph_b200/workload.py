"""Synthetic search / kNN workloads of the shapes BASELINE.json names, produced on the device with the product's own
kernels (data generator -> rg_knn -> rg_index_build_roargraph -> ground truth).  Used by bench.py and the -m gpu tests.

Nothing here touches the CPU checkers: they are only ever *compared against* these workloads by the callers.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

import roargraph_b200 as rg
from roargraph_b200 import synth

# the reference's canonical L_pq sweep (README.md:118, run_roargraph_search_test.sh)
README_L_SWEEP = [10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100, 110, 120, 130, 140, 150,
                  160, 170, 180, 190, 200, 220, 240, 260, 280, 300, 350, 400, 450, 500, 550, 600, 650, 700, 750, 800,
                  900, 1000, 1100, 1200, 1300, 1400, 1500, 1600, 1700, 1800, 1900, 2000]

_T0 = time.time()


def log(msg: str) -> None:
    print(f"[rg +{time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def _normalize_rows_(x):
    import torch
    n = torch.linalg.vector_norm(x, dim=1, keepdim=True).clamp_min(1e-30)
    x.div_(n)
    return x


def _pad8(x):
    import torch
    d = x.shape[1]
    pd = (d + 7) // 8 * 8
    if pd == d:
        return x.contiguous()
    out = torch.zeros((x.shape[0], pd), dtype=x.dtype, device=x.device)
    out[:, :d] = x
    return out


def recall_at_k(res_ids: np.ndarray, gt_ids: np.ndarray, k: int) -> float:
    """ComputeRecall of tests/test_search_roargraph.cpp:23-36, vectorised: |gt[:k] ∩ res[:k]| / (k nq)."""
    r = res_ids[:, :k].astype(np.int64)
    g = gt_ids[:, :k].astype(np.int64)
    hit = 0
    for j in range(k):  # membership test per ground-truth column, vectorised over the queries
        hit += int((r == g[:, j:j + 1]).any(axis=1).sum())
    return hit / float(k * res_ids.shape[0])


def bipartite_csr(knn_ids: np.ndarray, nd: int, M_pjbp: int):
    """Vectorised qbaseNNbipartite: query node nd + sq -> its first M_pjbp NNs except (every occurrence of) the top-1;
    base node top-1(sq) -> nd + sq, in ascending sq (the reference's single-thread order)."""
    nn = np.ascontiguousarray(knn_ids[:, :M_pjbp]).astype(np.uint32)
    nsq = nn.shape[0]
    tgt = nn[:, 0]
    keep = nn != tgt[:, None]
    q_deg = keep.sum(1)
    order = np.argsort(tgt, kind="stable")
    b_deg = np.bincount(tgt, minlength=nd)
    row_off = np.zeros(nd + nsq + 1, dtype=np.uint64)
    row_off[1:nd + 1] = np.cumsum(b_deg)
    row_off[nd + 1:] = row_off[nd] + np.cumsum(q_deg).astype(np.uint64)
    nbrs = np.concatenate([(order + nd).astype(np.uint32), nn[keep]])
    return row_off, nbrs


class SearchWorkload:
    """Device-resident shard of a synthetic cross-modal search problem + its RoarGraph index."""

    def __init__(self):
        self.index = None          # rg.RoarIndex over this rank's base rows
        self.d_base = None         # torch [nd_local, dim] (padded x8, normalised for cosine)
        self.d_queries = None      # torch [nq, dim] test queries, same treatment
        self.queries_host = None   # numpy copy of d_queries
        self.gt_ids = None         # numpy [nq, k] exact top-k over the FULL base (global ids)
        self.row_lo = 0            # first global base row of this shard
        self.nd_total = 0
        self.metric = "ip"
        self.k = 10
        self.d_train_sample = None # torch [<=151552, dim]: training queries kept for bench.py's brute-force kNN leg
        self.gen = None            # synth.TorchSynth: more queries of the same distribution on demand (make_queries)
        self.info = {}

    def make_queries(self, n: int):
        """n more queries from the generator's query distribution, treated like the test queries (padded to x8, normalised
        for cosine).  Deterministic given the seed and the call sequence, so every rank of a job draws the same rows."""
        q = _pad8(self.gen.queries(n))
        if self.metric == "cosine":
            _normalize_rows_(q)
        return q

    def exact_topk(self, d_q, k: int):
        """Exact top-k (global ids, uint32 numpy) of device-resident queries over the FULL base: the local brute-force kNN
        on one rank, the base-sharded kNN (all-gather + merge) when this workload is one shard of several."""
        import torch
        from roargraph_b200.sharded import ShardedKnn

        knn_metric = "l2" if self.metric == "l2" else "ip"
        sk = ShardedKnn(self.d_base, knn_metric, self.row_lo)
        try:
            ids, _ = sk.search(d_q, k)
            torch.cuda.synchronize(d_q.device)
            return ids.cpu().numpy().view(np.uint32)
        finally:
            sk.close()
            rg.knn_release_workspace()


BUILD_VERSION = 3  # bump when rg_index_build_roargraph's output changes (cached graphs of an older build must not be reused)


def _cache_key(**kw) -> str:
    s = repr(sorted(kw.items())) + f"abi{rg.lib().rg_abi_version()}build{BUILD_VERSION}"
    return hashlib.sha1(s.encode()).hexdigest()[:16]


def build_search_workload(nb: int, dim: int, metric: str, ntrain: int, nq: int, k: int = 10, M_sq: int = 100,
                          M_pjbp: int = 35, L_pjpq: int = 500, seed: int = 1234, device: int = 0, shard: int = 0,
                          n_shards: int = 1, cache_dir: str | None = None, n_rounds: int = 0) -> SearchWorkload:
    """base nb x dim (+ ntrain training queries, nq test queries) from the seeded generator; this
    rank keeps base rows [shard * nb / n_shards, (shard + 1) * nb / n_shards) and builds its own RoarGraph on them from
    the training queries' exact M_sq-NN lists within the shard (the sharding contract of SURVEY.md 8(e))."""
    import torch

    dev = torch.device("cuda", device)
    w = SearchWorkload()
    w.metric, w.k, w.nd_total = metric, k, nb
    t0 = time.time()
    base, train, test, w.gen = synth.make_torch(nb, ntrain, nq, dim, dev, seed=seed, return_gen=True)
    base, train, test = _pad8(base), _pad8(train), _pad8(test)
    if metric == "cosine":  # the reference normalises base rows at load and queries in the CLI
        _normalize_rows_(base)
        _normalize_rows_(train)
        _normalize_rows_(test)
    torch.cuda.synchronize(dev)
    log(f"synthetic data on device: base {tuple(base.shape)}, train {tuple(train.shape)}, test {tuple(test.shape)} "
        f"({time.time() - t0:.1f}s)")
    knn_metric = "l2" if metric == "l2" else "ip"

    # exact ground truth of the test queries over the FULL base (every rank computes the same thing)
    t0 = time.time()
    gt_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    gt_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    rg.knn_dev(base, test, k, knn_metric, gt_i, gt_d)
    torch.cuda.synchronize(dev)
    w.gt_ids = gt_i.cpu().numpy().astype(np.uint32)
    w.info["gt_s"] = round(time.time() - t0, 2)
    del gt_i, gt_d

    lo = shard * nb // n_shards
    hi = (shard + 1) * nb // n_shards
    w.row_lo = lo
    if n_shards > 1:
        local = base[lo:hi].clone()
        del base
        base = local
        torch.cuda.empty_cache()
    w.d_base = base
    w.d_queries = test
    w.queries_host = test.cpu().numpy()
    nd = hi - lo

    w.index = rg.RoarIndex(base, "l2" if metric == "l2" else "ip")
    key = _cache_key(nb=nb, dim=dim, metric=metric, ntrain=ntrain, M_sq=M_sq, M_pjbp=M_pjbp, L_pjpq=L_pjpq, seed=seed,
                     shard=shard, n_shards=n_shards, n_rounds=n_rounds)
    cpath = os.path.join(cache_dir, f"graph_{key}.npz") if cache_dir else None
    if cpath and os.path.exists(cpath):
        t0 = time.time()
        z = np.load(cpath)
        w.index.set_graph(int(z["ep"]), z["row_off"], z["nbrs"])
        w.info["graph_from_cache_s"] = round(time.time() - t0, 2)
        log(f"RoarGraph index loaded from {cpath} ({time.time() - t0:.1f}s)")
    else:
        t0 = time.time()
        kn_i = torch.empty((ntrain, M_sq), dtype=torch.int32, device=dev)
        kn_d = torch.empty((ntrain, M_sq), dtype=torch.float32, device=dev)
        rg.knn_dev(base, train, M_sq, knn_metric, kn_i, kn_d)
        torch.cuda.synchronize(dev)
        w.info["train_knn_s"] = round(time.time() - t0, 2)
        w.info["train_knn_pairs_per_s"] = float(ntrain) * nd / max(time.time() - t0, 1e-9)
        log(f"training-query kNN {ntrain} x {nd}, K={M_sq}: {time.time() - t0:.1f}s")
        del kn_d
        t0 = time.time()
        w.index.build_roargraph(kn_i, M_sq, M_pjbp, L_pjpq, n_rounds)
        w.info["build_s"] = round(time.time() - t0, 2)
        del kn_i
        g = w.index.graph_size()
        log(f"BuildRoarGraph on device: {time.time() - t0:.1f}s, avg degree {g['nnz'] / nd:.1f}, max {g['max_degree']}, "
            f"ep {g['ep']}")
        if cpath:
            os.makedirs(cache_dir, exist_ok=True)
            ep, row_off, nbrs = w.index.get_graph()
            tmp = cpath + f".tmp{os.getpid()}.npz"
            np.savez(tmp, ep=np.uint32(ep), row_off=row_off, nbrs=nbrs)
            os.replace(tmp, cpath)
    w.d_train_sample = train[:min(ntrain, 151552)].clone()
    del train
    rg.knn_release_workspace()  # the kNN scratch (bf16 base copy, candidate lists) is not needed while searching
    torch.cuda.empty_cache()
    g = w.index.graph_size()
    w.info.update(nd_local=nd, avg_degree=round(g["nnz"] / nd, 2), max_degree=g["max_degree"], ep=g["ep"])
    return w
