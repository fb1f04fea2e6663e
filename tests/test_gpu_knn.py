"""GPU parity tests for the brute-force kNN (exact_knn) through the C-ABI, against the oracle and the reference's
compute_groundtruth output stored in tests/golden."""
import os

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


def test_knn_seed_pass_changes_nothing_observable(knn_mode, monkeypatch):
    """The tau seed pass (bf16 group maxima of the first 65 536 rows -> an upper bound of every query's K-th best distance)
    only removes work: ids and distance bits equal those of the unseeded ramp, and both equal the resident-base entry
    (rg_knn_base_*) fed the queries in two batches."""
    if knn_mode != "tc":
        pytest.skip("the seed pass belongs to the 2-CTA tensor-core filter")
    import torch
    rng = np.random.default_rng(5)
    nb, nq, dim = 300_017, 2500, 200
    cent = rng.standard_normal((64, dim)).astype(np.float32)
    base = (cent[rng.integers(0, 64, nb)] + 0.7 * rng.standard_normal((nb, dim))).astype(np.float32)
    q = (cent[rng.integers(0, 64, nq)] + 0.9 * rng.standard_normal((nq, dim))).astype(np.float32)
    q[7] = 0.0  # a zero query: every score ties, the seed bound must not drop everything
    for metric, K in (("ip", 100), ("l2", 10), ("cosine", 100)):
        monkeypatch.setenv("RG_KNN_SEED", "0")
        ui, ud = rg.knn(base, q, K, metric)
        monkeypatch.delenv("RG_KNN_SEED")
        si, sd = rg.knn(base, q, K, metric)
        assert (ui == si).all() and (ud.view(np.uint32) == sd.view(np.uint32)).all(), metric
        kb = rg.KnnBase(base, metric)
        d_q = torch.from_numpy(q).cuda()
        d_i = torch.empty((nq, K), dtype=torch.int32, device="cuda")
        d_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
        kb.search_dev(d_q[:1000], K, d_i[:1000], d_d[:1000], id_offset=5)
        kb.search_dev(d_q[1000:], K, d_i[1000:], d_d[1000:], id_offset=5)
        torch.cuda.synchronize()
        assert (d_i.cpu().numpy() == si + 5).all() and (d_d.cpu().numpy().view(np.uint32) == sd.view(np.uint32)).all(), metric
        kb.close()


@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_knn_vs_mkl_reference_at_scale(knn_mode, metric, tmp_path):
    """200 000 x 200 000 x 200, K = 100 against the reference's own compute_groundtruth (MKL SGEMM, compiled unmodified,
    compute_groundtruth.cpp:126-248) at the norms and K of the real workload — where the bf16 filter threshold
    (rg_knn_tc.cu: theta, kErrRel) and the tau seed are exercised for real: distances to 1e-4 relative, ids equal except
    inside near-ties of the reference's own list."""
    import os
    from oracle.oracle import HERE as ORACLE_DIR, run_ref_compute_groundtruth
    from roargraph_b200 import formats as F, synth
    if knn_mode != "tc":
        pytest.skip("one filter is enough for the 80-second MKL run")
    if not os.path.exists(os.path.join(ORACLE_DIR, "_ref", "ref_compute_groundtruth")):
        pytest.skip("oracle/_ref/ref_compute_groundtruth not available")
    nb = nq = 200_000
    base, train, _ = synth.make_numpy(nb, nq, 1, 200, seed=4242)
    bf, qf, gf = (str(tmp_path / n) for n in ("base.fbin", "train.fbin", "train.gt"))
    F.write_fbin(bf, base)
    F.write_fbin(qf, train)
    run_ref_compute_groundtruth(bf, qf, gf, 100, "mips" if metric == "ip" else "l2")
    ref_i, ref_d = F.read_truthset(gf)
    if metric == "ip":
        ref_d = -ref_d  # the tool stores +IP for mips (compute_groundtruth.cpp:438-439)
    ids, dists = rg.knn(base, train, 100, metric)
    exact = check_knn(ids.astype(np.uint32), dists, ref_i, ref_d)
    assert exact > 0.99, exact  # MKL-vs-MKL already differs on ~0.003 % of ids from ties
    recall = np.mean([len(set(a) & set(b)) for a, b in zip(ids[::97], ref_i[::97].astype(np.int32))]) / 100.0
    assert recall > 0.9999, recall


def test_knn_shared_tau_across_shards_is_exact(knn_mode):
    """The sharded-base protocol on one GPU: two shards of a 400 000-row base each offer a bound of their ceil(K/2)-th best seed
    row (rg_knn_base_seed_dev), the maximum is put in force on both (rg_knn_base_search_seeded_dev), the per-shard lists —
    now possibly shorter than K — are packed, merged, and must equal the unsharded answer bit for bit."""
    if knn_mode != "tc":
        pytest.skip("the seed pass belongs to the 2-CTA tensor-core filter")
    import torch
    rng = np.random.default_rng(8)
    nb, nq, dim, K = 400_000, 3000, 200, 100
    cent = rng.standard_normal((64, dim)).astype(np.float32)
    base = (cent[rng.integers(0, 64, nb)] + 0.7 * rng.standard_normal((nb, dim))).astype(np.float32)
    q = (cent[rng.integers(0, 64, nq)] + 0.9 * rng.standard_normal((nq, dim))).astype(np.float32)
    d_base, d_q = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
    for metric in ("ip", "l2"):
        want_i, want_d = rg.knn(base, q, K, metric)
        G, lo = 2, [0, nb // 2, nb]
        shards = [rg.KnnBase(d_base[lo[g]:lo[g + 1]], metric) for g in range(G)]
        taus = []
        for kb in shards:
            t = torch.empty(nq, dtype=torch.float32, device="cuda")
            kb.seed_dev(d_q, K, -(-K // G), t)
            taus.append(t)
        tau = torch.maximum(taus[0], taus[1])
        assert torch.isfinite(tau).all()
        packed = torch.empty((G, nq, K), dtype=torch.int64, device="cuda")
        short = 0
        for g, kb in enumerate(shards):
            li = torch.empty((nq, K), dtype=torch.int32, device="cuda")
            ld = torch.empty((nq, K), dtype=torch.float32, device="cuda")
            # bound_rows + RG_KNN_WHOLE_SHARD=1: with a generous claim the whole shard goes through ONE filter + merge pair
            # (and falls back to the growing schedule if a list overflows); both schedules must give the same lists
            if g == 1:
                os.environ["RG_KNN_WHOLE_SHARD"] = "1"
            try:
                kb.search_seeded_dev(d_q, K, tau, li, ld, bound_rows=(0 if g == 0 else 8 * nb))
            finally:
                os.environ.pop("RG_KNN_WHOLE_SHARD", None)
            short += int((li < 0).sum().item())
            rg.topk_pack_dev(ld, li, lo[g], packed[g])
        out_i = torch.empty((nq, K), dtype=torch.int32, device="cuda")
        out_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
        rg.topk_merge_packed_dev(packed, out_d, out_i)
        torch.cuda.synchronize()
        # (with 8 shards the shared bound cuts most per-shard lists below K — `short` counts the -1 ids; with 2 shards of
        # 200 000 rows each still holds ~150 rows under it, so the lists usually stay full here)
        assert (out_i.cpu().numpy() == want_i).all(), metric
        assert (out_d.cpu().numpy().view(np.uint32) == want_d.view(np.uint32)).all(), metric
        for kb in shards:
            kb.close()


def test_knn_bound_resync_inside_the_shard_is_exact(knn_mode):
    """rg_knn_base_search_resync_dev on one GPU: two shards, the seed bound as above, plus ONE re-sync inside each shard whose
    hook plays the other shard's part — it raises the published rank-ceil(K/2) distances to the maximum with the OTHER shard's
    final rank-ceil(K/2) distance (a value that shard certainly holds ceil(K/2) rows under).  The merged lists must equal the
    unsharded answer bit for bit, the hook must run exactly once per call, and the tightened bound must really cut work."""
    if knn_mode != "tc":
        pytest.skip("the seed pass belongs to the 2-CTA tensor-core filter")
    import torch
    rng = np.random.default_rng(18)
    nb, nq, dim, K = 600_000, 2500, 200, 100
    cent = rng.standard_normal((64, dim)).astype(np.float32)
    base = (cent[rng.integers(0, 64, nb)] + 0.7 * rng.standard_normal((nb, dim))).astype(np.float32)
    q = (cent[rng.integers(0, 64, nq)] + 0.9 * rng.standard_normal((nq, dim))).astype(np.float32)
    d_base, d_q = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
    G, lo, r = 2, [0, nb // 2 + 1280, nb], 50
    for metric in ("ip", "l2"):
        want_i, want_d = rg.knn(base, q, K, metric)
        shards = [rg.KnnBase(d_base[lo[g]:lo[g + 1]], metric) for g in range(G)]
        taus, finals = [], []
        for kb in shards:
            t = torch.empty(nq, dtype=torch.float32, device="cuda")
            kb.seed_dev(d_q, K, r, t)
            taus.append(t)
            li = torch.empty((nq, K), dtype=torch.int32, device="cuda")
            ld = torch.empty((nq, K), dtype=torch.float32, device="cuda")
            kb.search_dev(d_q, K, li, ld)
            finals.append(ld[:, r - 1].contiguous())
        tau = torch.maximum(taus[0], taus[1])
        packed = torch.empty((G, nq, K), dtype=torch.int64, device="cuda")
        for resync_rows in (0, (nb // 2) // 5, 10 * nb):      # after the first chunk / a fifth of the shard / after the last chunk
            calls, tight = [], []
            for g, kb in enumerate(shards):
                li = torch.empty((nq, K), dtype=torch.int32, device="cuda")
                ld = torch.empty((nq, K), dtype=torch.float32, device="cuda")
                bound = torch.full((nq,), -1.0, dtype=torch.float32, device="cuda")

                def reduce_max(off, cnt, bound=bound, g=g):
                    calls.append((g, off, cnt))
                    mine = bound[off:off + cnt]
                    tight.append(float((torch.maximum(mine, finals[1 - g][off:off + cnt]) < tau[off:off + cnt]).float().mean()))
                    torch.maximum(mine, finals[1 - g][off:off + cnt], out=mine)

                kb.search_resync_dev(d_q, K, tau, li, ld, r, resync_rows, bound, reduce_max)
                rg.topk_pack_dev(ld, li, lo[g], packed[g])
            assert calls == [(0, 0, nq), (1, 0, nq)], calls
            out_i = torch.empty((nq, K), dtype=torch.int32, device="cuda")
            out_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
            rg.topk_merge_packed_dev(packed, out_d, out_i)
            torch.cuda.synchronize()
            assert (out_i.cpu().numpy() == want_i).all(), (metric, resync_rows)
            assert (out_d.cpu().numpy().view(np.uint32) == want_d.view(np.uint32)).all(), (metric, resync_rows)
            if resync_rows == (nb // 2) // 5:
                assert min(tight) > 0.5, tight   # the re-synced bound is below the seed bound for most queries
        # a hook that fails must surface as an error, not as a hang or a wrong answer
        with pytest.raises(ZeroDivisionError):
            shards[0].search_resync_dev(d_q, K, tau, li, ld, r, 0, bound, lambda off, cnt: 1 // 0)
        for kb in shards:
            kb.close()


def test_knn_merge_by_rank_equals_full_sort(knn_mode, monkeypatch):
    """The re-rank kernel's two list updates — sort the new entries and merge with the sorted list by rank (default), or sort
    everything (RG_KNN_MERGE_RANK=0) — must leave identical lists: ids and distance bits, short lists (K close to the rows
    seen in the first chunks) and long ones."""
    if knn_mode != "tc":
        pytest.skip("the fast merge kernel belongs to the tensor-core pipeline")
    import torch
    rng = np.random.default_rng(21)
    nb, nq, dim = 300_000, 3000, 200
    base = rng.standard_normal((nb, dim)).astype(np.float32) * (1.0 + 0.3 * rng.random((nb, 1), dtype=np.float32))
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    base[1000:1040] = base[5]                 # exact duplicates: equal distances, ordered by id
    d_base, d_q = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
    for metric in ("ip", "l2"):
        for K in (1, 10, 100, 500):
            got = []
            for flag in ("1", "0"):
                monkeypatch.setenv("RG_KNN_MERGE_RANK", flag)
                ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
                dd = torch.empty((nq, K), dtype=torch.float32, device="cuda")
                rg.knn_dev(d_base, d_q, K, metric, ids, dd)
                torch.cuda.synchronize()
                got.append((ids.cpu().numpy(), dd.cpu().numpy().view(np.uint32)))
            assert (got[0][0] == got[1][0]).all(), (metric, K)
            assert (got[0][1] == got[1][1]).all(), (metric, K)
            assert (got[0][0] >= 0).all()
