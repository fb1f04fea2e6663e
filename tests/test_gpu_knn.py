"""GPU parity tests for the brute-force kNN (exact_knn) through the C-ABI, against the oracle and the reference's
compute_groundtruth output stored in tests/golden."""
import numpy as np
import pytest

import roargraph_b200 as rg
from test_oracle_golden import check_knn

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tc", "tc1", "tc2", "simt"], autouse=True)
def knn_mode(request, monkeypatch):
    """Every test runs through every candidate filter: the 2-CTA tcgen05 bf16 kernel (default), the single-CTA tcgen05
    kernels with one / two resident query tiles, and the fp32 SIMT tile."""
    monkeypatch.setenv("RG_KNN_MODE", "simt" if request.param == "simt" else "tc")
    if request.param in ("tc1", "tc2"):
        monkeypatch.setenv("RG_KNN_QTILES", request.param[2])
    else:
        monkeypatch.delenv("RG_KNN_QTILES", raising=False)
    return request.param


def test_knn_golden(golden):
    K = golden.z["gt_ids"].shape[1]
    ids, dists = rg.knn(golden.base_raw, golden.test_raw, K, golden.metric)
    ref_d = golden.z["gt_dists"]
    if golden.metric == "ip":
        ref_d = -ref_d
    assert check_knn(ids.astype(np.uint32), dists, golden.z["gt_ids"], ref_d) > 0.9
    assert (np.diff(dists, axis=1) >= 0).all()


@pytest.mark.parametrize("nb,nq,dim,K,metric", [(5000, 77, 200, 100, "ip"), (3001, 130, 40, 1, "l2"),
                                                 (9000, 257, 104, 10, "cosine"), (300, 5, 8, 300, "ip"),
                                                 (20000, 64, 512, 100, "l2")])
def test_knn_vs_oracle(oracle, nb, nq, dim, K, metric):
    rng = np.random.default_rng(nb)
    base = rng.standard_normal((nb, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ids, dists = rg.knn(base, q, K, metric)
    oi, od = oracle.exact_knn(base, q, K, metric)
    check_knn(ids.astype(np.uint32), dists, oi.astype(np.uint32), od)
    # id_offset (start_id of a base part, compute_groundtruth.cpp:413)
    ids2, _ = rg.knn(base, q[:3], K, metric, id_offset=1000)
    assert (ids2 == ids[:3] + 1000).all()


def test_knn_sorted_base_triggers_overflow_fallback(oracle):
    """Base rows ordered from worst to best for every query: every chunk beats tau everywhere, the candidate lists
    overflow and the safe chunking mode must give the same answer."""
    rng = np.random.default_rng(9)
    dim, nb = 16, 40000
    direction = rng.standard_normal(dim).astype(np.float32)
    scale = np.linspace(0.01, 3.0, nb, dtype=np.float32)[:, None]
    base = (scale * direction[None, :] + 0.001 * rng.standard_normal((nb, dim))).astype(np.float32)
    q = (direction[None, :] * np.array([[1.0], [2.0], [0.5]], dtype=np.float32)).astype(np.float32)
    ids, dists = rg.knn(base, q, 10, "ip")
    oi, od = oracle.exact_knn(base, q, 10, "ip")
    check_knn(ids.astype(np.uint32), dists, oi.astype(np.uint32), od)


def test_topk_merge():
    import torch
    rng = np.random.default_rng(2)
    G, nq, K = 4, 50, 10
    d = np.sort(rng.standard_normal((G, nq, K)).astype(np.float32), axis=2)
    i = rng.permutation(G * nq * K).astype(np.uint32).reshape(G, nq, K)
    dd, di = torch.from_numpy(d).cuda(), torch.from_numpy(i.view(np.int32)).cuda()
    od = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    oi = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    rg.topk_merge_dev(dd, di, od, oi)
    torch.cuda.synchronize()
    alld = d.transpose(1, 0, 2).reshape(nq, G * K)
    alli = i.transpose(1, 0, 2).reshape(nq, G * K)
    order = np.lexsort((alli, alld), axis=1)[:, :K]
    assert (np.take_along_axis(alld, order, 1) == od.cpu().numpy()).all()
    assert (np.take_along_axis(alli, order, 1) == oi.cpu().numpy().view(np.uint32)).all()


def test_knn_tensor_core_matches_simt_at_scale(monkeypatch):
    """300k x 200d base, 3000 queries, K=100: the bf16 tensor-core filter + fp32 re-rank must return the fp32 answer
    (same ids except inside near-ties, distances to 1e-5 relative), for IP and L2, incl. partial query / base tiles."""
    rng = np.random.default_rng(11)
    nb, nq, dim, K = 300_017, 3001, 200, 100
    cent = rng.standard_normal((64, dim)).astype(np.float32)
    base = (cent[rng.integers(0, 64, nb)] + 0.7 * rng.standard_normal((nb, dim))).astype(np.float32)
    q = (cent[rng.integers(0, 64, nq)] + 0.9 * rng.standard_normal((nq, dim))).astype(np.float32)
    for metric in ("ip", "l2"):
        monkeypatch.setenv("RG_KNN_MODE", "simt")
        si, sd = rg.knn(base, q, K, metric)
        monkeypatch.setenv("RG_KNN_MODE", "tc")
        ti, td = rg.knn(base, q, K, metric)
        np.testing.assert_allclose(td, sd, rtol=2e-5, atol=1e-4)
        same = (ti == si)
        assert same.mean() > 0.999, same.mean()
        # wherever ids differ the two distances must be a near-tie
        bad = ~same
        assert np.all(np.abs(td[bad] - sd[bad]) <= 1e-4 + 2e-5 * np.abs(sd[bad]))
        assert (np.diff(td, axis=1) >= 0).all()


def test_knn_host_entry_streams_query_batches(oracle, monkeypatch):
    """rg_knn keeps the base resident and streams the queries in batches (ragged last batch): same answer as one batch."""
    rng = np.random.default_rng(11)
    base = rng.standard_normal((4000, 72)).astype(np.float32)
    q = rng.standard_normal((333, 72)).astype(np.float32)
    whole_i, whole_d = rg.knn(base, q, 20, "l2")
    monkeypatch.setenv("RG_KNN_HOST_BATCH", "100")
    part_i, part_d = rg.knn(base, q, 20, "l2")
    assert (whole_i == part_i).all() and (whole_d.view(np.uint32) == part_d.view(np.uint32)).all()
    oi, od = oracle.exact_knn(base, q, 20, "l2")
    check_knn(part_i.astype(np.uint32), part_d, oi.astype(np.uint32), od)
