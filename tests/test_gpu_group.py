"""GPU tests of the single-process multi-GPU entry points (rg_group_* / rg_sharded_* of include/rg_b200.h) through the C-ABI:
base rows sharded over the GPUs of one process, per-shard RoarGraph, ONE packed all-gather + merge — against the CPU oracle
run on every shard's own index and merged on the host by (distance, id), the oracle SURVEY.md 8(e) names for this path.
With one visible GPU the group degenerates to a single shard (same code path without NCCL); with two or more (gpurun
--gpus 2) the exchange runs over NCCL."""
import numpy as np
import pytest

import roargraph_b200 as rg
from roargraph_b200 import formats as F, synth
from test_oracle_golden import check_knn

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("G", [1, 2, 4])
def test_group_sharded_search_equals_per_shard_oracle_merge(oracle, G):
    if _n_gpus() < G:
        pytest.skip(f"needs {G} GPUs")
    nb, dim, k = 24000, 104, 10
    base, train, test = synth.make_numpy(nb, 6000, 200, dim, seed=21)
    base, train, test = F.pad_rows(base), F.pad_rows(train), F.pad_rows(test)
    grp = rg.Group(G)
    si = rg.GroupShardedIndex(grp, base, "l2")
    si.build_roargraph(train, 64, 24, 200)
    shards = []
    for s in range(G):
        idx, lo, n = si.shard(s)
        ep, row_off, nbrs = idx.get_graph()
        assert lo == s * nb // G and n == (s + 1) * nb // G - lo and len(row_off) == n + 1
        shards.append((lo, n, ep, row_off, nbrs))
    for L in (10, 40, 150):
        ids, dists = si.search(test, k, L)
        parts_i, parts_d = [], []
        for lo, n, ep, row_off, nbrs in shards:
            r = oracle.search_roargraph(base[lo:lo + n], row_off, nbrs, ep, "l2", test, k, L)
            assert r["nfail"] == 0
            parts_i.append(r["ids"].astype(np.int64) + lo)
            parts_d.append(r["dists"])
        ai, ad = np.concatenate(parts_i, 1), np.concatenate(parts_d, 1)
        order = np.lexsort((ai, ad), axis=1)[:, :k]
        assert (np.take_along_axis(ai, order, 1) == ids).all(), f"G={G} L={L}"
        assert (np.take_along_axis(ad, order, 1).view(np.uint32) == dists.view(np.uint32)).all()
    # a batch smaller than the group and a ragged one
    for nq in (1, 7):
        ids1, _ = si.search(test[:nq], k, 40)
        assert (ids1 == si.search(test, k, 40)[0][:nq]).all()
    si.close()
    grp.close()


@pytest.mark.parametrize("G", [1, 2, 4])
def test_group_knn_equals_oracle(oracle, G, monkeypatch):
    if _n_gpus() < G:
        pytest.skip(f"needs {G} GPUs")
    rng = np.random.default_rng(3)
    base = rng.standard_normal((9001, 72)).astype(np.float32)
    q = rng.standard_normal((333, 72)).astype(np.float32)
    grp = rg.Group(G)
    for metric, K in (("l2", 20), ("ip", 100)):
        oi, od = oracle.exact_knn(base, q, K, metric)
        ids, dists = grp.knn(base, q, K, metric)
        check_knn(ids.astype(np.uint32), dists, oi.astype(np.uint32), od)
        monkeypatch.setenv("RG_KNN_HOST_BATCH", "100")  # several ragged batches through the exchange
        ids2, dists2 = grp.knn(base, q, K, metric)
        monkeypatch.delenv("RG_KNN_HOST_BATCH")
        assert (ids2 == ids).all() and (dists2.view(np.uint32) == dists.view(np.uint32)).all()
    grp.close()


@pytest.mark.parametrize("G", [2, 4])
def test_group_knn_shared_bound_at_scale_equals_single_gpu(G, monkeypatch):
    """Shards large enough for the tensor-core seed pass (the shared bound really cuts the per-shard lists), with and without
    the re-sync of the bound inside the shard (RG_KNN_RESYNC=1), several host batches: ids and distance bits of the single-GPU
    kNN over the whole base."""
    if _n_gpus() < G:
        pytest.skip(f"needs {G} GPUs")
    rng = np.random.default_rng(5)
    nb, nq, dim, K = 520_000, 2500, 200, 100
    cent = rng.standard_normal((64, dim)).astype(np.float32)
    base = (cent[rng.integers(0, 64, nb)] + 0.7 * rng.standard_normal((nb, dim))).astype(np.float32)
    q = (cent[rng.integers(0, 64, nq)] + 0.9 * rng.standard_normal((nq, dim))).astype(np.float32)
    grp = rg.Group(G)
    monkeypatch.setenv("RG_KNN_HOST_BATCH", "1024")
    for metric in ("ip", "l2"):
        want_i, want_d = rg.knn(base, q, K, metric)
        for resync in ("0", "1"):
            monkeypatch.setenv("RG_KNN_RESYNC", resync)
            ids, dists = grp.knn(base, q, K, metric)
            assert (ids == want_i).all(), (metric, resync)
            assert (dists.view(np.uint32) == want_d.view(np.uint32)).all(), (metric, resync)
    grp.close()
