"""GPU parity tests for the beam-search kernels, through the C-ABI (librg_b200.so) against the CPU oracle."""
import os

import numpy as np
import pytest

import roargraph_b200 as rg
from gpu_helpers import assert_same_search, bipartite_from_knn, bits, build_reference_case, random_graph
from oracle.oracle import have_detrand_reference, have_reference

pytestmark = pytest.mark.gpu


def test_golden_parity(oracle, golden):
    """Committed reference outputs: ids, distance bits, cmps, hops identical at every stored L_pq."""
    base, queries = golden.normalised(oracle)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    for L in golden.Ls:
        want = {k: golden.z[f"{k}_L{L}"] for k in ("ids", "dists", "cmps", "hops")}
        assert_same_search(idx.search(queries, golden.k, L), want, f"{golden.name} L={L}")
    idx.close()


@pytest.mark.parametrize("L", [10, 11, 33, 64, 257, 2000])
def test_oracle_parity_various_L(oracle, golden, L):
    base, queries = golden.normalised(oracle)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    for k in (1, 10):
        want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries, k, L)
        assert_same_search(idx.search(queries, k, L), want, f"{golden.name} L={L} k={k}")
    # k == L (on these 2000-4000-node graphs L=2000 can exceed the reachable set: then both sides must fail alike)
    want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries[:7], L, L)
    if want["nfail"]:
        with pytest.raises(rg.RgError) as e:
            idx.search(queries[:7], L, L)
        assert e.value.code == rg.RG_ENOTENOUGH
    else:
        assert_same_search(idx.search(queries[:7], L, L), want, "k==L")
    idx.close()


@pytest.mark.parametrize("dim,metric,deg", [(8, "ip", (0, 5)), (16, "l2", (1, 40)), (200, "ip", (20, 70)),
                                            (512, "ip", (5, 33)), (24, "l2", (90, 150))])
def test_random_graphs_with_duplicates_selfloops_and_big_degrees(oracle, dim, metric, deg):
    """Adversarial graphs: repeated ids inside rows, self loops, empty rows, rows longer than one 96-id pass."""
    rng = np.random.default_rng(dim)
    nd = 3000
    base = rng.standard_normal((nd, dim)).astype(np.float32)
    queries = rng.standard_normal((37, dim)).astype(np.float32)
    row_off, nbrs = random_graph(rng, nd, deg[0], deg[1], dup_frac=0.3, self_loops=True, zero_frac=0.05)
    ep = 5
    idx = rg.RoarIndex(base, metric)
    idx.set_graph(ep, row_off, nbrs)
    for L in (10, 100):
        want = oracle.search_roargraph(base, row_off, nbrs, ep, metric, queries, 1, L)
        ok = want["status"] == 0
        if not ok.all():
            with pytest.raises(rg.RgError) as e:
                idx.search(queries, 1, L)
            assert e.value.code == rg.RG_ENOTENOUGH
            continue
        assert_same_search(idx.search(queries, 1, L), want, f"random dim={dim} L={L}")
    idx.close()


def test_errors(oracle):
    rng = np.random.default_rng(0)
    base = rng.standard_normal((100, 16)).astype(np.float32)
    q = rng.standard_normal((3, 16)).astype(np.float32)
    with pytest.raises(rg.RgError) as e:
        rg.RoarIndex(rng.standard_normal((10, 12)).astype(np.float32), "ip")  # dim not padded to x8
    assert e.value.code == rg.RG_EINVAL
    idx = rg.RoarIndex(base, "ip")
    with pytest.raises(rg.RgError) as e:
        idx.search(q, 1, 10)  # graph not set
    assert e.value.code == rg.RG_ESTATE
    row_off = np.zeros(101, dtype=np.uint64)
    idx.set_graph(0, row_off, np.zeros(0, np.uint32))
    with pytest.raises(rg.RgError) as e:
        idx.search(q, 11, 10)  # k > L_pq (CLI: "L_pq must greater or equal than k")
    assert e.value.code == rg.RG_EINVAL
    with pytest.raises(rg.RgError) as e:
        idx.search(q, 5, 10)  # edge-less graph: only the entry point is found
    assert e.value.code == rg.RG_ENOTENOUGH and "not enough results: 1" in str(e.value)
    ids, dists, cmps, hops = idx.search(q, 1, 10)
    assert (ids == 0).all() and (cmps == 0).all() and (hops == 1).all()
    with pytest.raises(rg.RgError):
        idx.set_graph(100, row_off, np.zeros(0, np.uint32))  # entry point out of range
    idx.close()


def test_device_pointer_entry_and_stats(oracle, golden):
    import torch
    base, queries = golden.normalised(oracle)
    d_base = torch.from_numpy(base).cuda()
    idx = rg.RoarIndex(d_base, golden.metric)
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    L, k = 32, golden.k
    d_q = torch.from_numpy(queries).cuda()
    nq = d_q.shape[0]
    d_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    d_dists = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    d_cmps = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_hops = torch.empty(nq, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    idx.search_dev(d_q, k, L, d_ids, d_dists, d_cmps, d_hops, stream=s)
    idx.check(s)
    want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries, k, L)
    got = (d_ids.cpu().numpy().view(np.uint32), d_dists.cpu().numpy(), d_cmps.cpu().numpy().view(np.uint32),
           d_hops.cpu().numpy().view(np.uint32))
    assert_same_search(got, want, "device-pointer entry")
    st = idx.last_stats()
    assert st["cmps"] == int(want["cmps"].sum()) and st["hops"] == int(want["hops"].sum())
    assert st["adj_words"] == int(want["adj_words"].sum())
    idx.close()


@pytest.mark.parametrize("where", ["hash", "smem"])
@pytest.mark.parametrize("bbits", [1, 3, 6])
def test_visited_table_overflow_migrates_to_tags(oracle, golden, monkeypatch, bbits, where):
    """The hashed visited table (L2-resident or in shared memory) forced tiny (2^bbits buckets of 16 ids): bucket pairs
    overflow in mid-query and the query migrates to the tag array; ids, distance bits, cmps and hops must not change."""
    monkeypatch.setenv("RG_VISITED", where)
    monkeypatch.setenv("RG_VTAB_BBITS", str(bbits))
    base, queries = golden.normalised(oracle)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    for L in (10, 64, 257):
        want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries, 10, L)
        assert_same_search(idx.search(queries, 10, L), want, f"{golden.name} tiny visited table bbits={bbits} L={L}")
    idx.close()


@pytest.mark.parametrize("where", ["hash", "smem"])
def test_visited_table_default_size_matches_oracle(oracle, golden, monkeypatch, where):
    """The table at its planned size (no forced overflow): every visited check is answered by the table alone."""
    monkeypatch.setenv("RG_VISITED", where)
    base, queries = golden.normalised(oracle)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    for L in (10, 40, 130, 300):
        want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries, 10, L)
        assert_same_search(idx.search(queries, 10, L), want, f"{golden.name} visited={where} L={L}")
    idx.close()


@pytest.mark.parametrize("visited", ["tags", "hash", "smem"])
def test_visited_epoch_wrap(oracle, monkeypatch, visited):
    """> 255 queries through the same warp slot: the 8-bit visited epoch wraps and the tag array is cleared
    (VisitedList::reset, visited_list_pool.h:20-26).  "hash": a 2-bucket visited table, so every query migrates to the
    tags and bumps the epoch on the way."""
    from conftest import GoldenCase
    g = GoldenCase("ip_d40")
    monkeypatch.setenv("RG_VISITED", visited)
    if visited != "tags":
        monkeypatch.setenv("RG_VTAB_BBITS", "1")
    monkeypatch.setenv("RG_CTAS_PER_SM", "1")
    monkeypatch.setenv("RG_WARPS_PER_CTA", "1")
    base, queries = g.base, g.queries
    reps = (rg.sm_count() * 300) // queries.shape[0] + 1
    big = np.ascontiguousarray(np.tile(queries, (reps, 1)))
    idx = rg.RoarIndex(base, "ip")
    idx.set_graph(g.ep, g.row_off, g.nbrs)
    ids, dists, cmps, hops = idx.search(big, g.k, 10)
    want = {k: np.tile(g.z[f"{k}_L10"], (reps, 1) if g.z[f"{k}_L10"].ndim == 2 else reps) for k in ("ids", "dists", "cmps", "hops")}
    assert_same_search((ids, dists, cmps, hops), want, "epoch wrap")
    idx.close()


def test_bipartite_parity(oracle, golden):
    """SearchBipartiteGraph with injected entry points vs the oracle restatement (src/index_bipartite.cpp:282-356)."""
    base, queries = golden.normalised(oracle)
    nd = base.shape[0]
    row_off, nbrs = bipartite_from_knn(golden.z["train_knn_ids"], nd, 12)
    rng = np.random.default_rng(1)
    init = rng.integers(0, nd, (queries.shape[0], 10)).astype(np.uint32)
    init[0, 1] = init[0, 0]  # repeated entry point
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_bipartite(row_off, nbrs)
    for L in (10, 50, 300):
        want = oracle.search_bipartite(base, row_off, nbrs, init, golden.metric, queries, golden.k, L)
        assert want["nfail"] == 0
        assert_same_search(idx.search_bipartite(queries, golden.k, L, init_ids=init), want, f"bipartite L={L}")
    # seeded entry points: deterministic, in range, and equal to running with those ids injected
    # (k=3: ten random entry points may repeat, and base nodes that are nobody's top-1 have no out-edges)
    a = idx.search_bipartite(queries, 3, 50, seed=42)
    b = idx.search_bipartite(queries, 3, 50, seed=42)
    assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all()
    idx.close()


def test_bipartite_golden_parity(oracle, golden):
    """Committed outputs of the compiled reference's SearchBipartiteGraph (tests/golden/bipartite_*.npz)."""
    from conftest import BipartiteGolden
    b = BipartiteGolden(golden.name)
    base, queries = golden.normalised(oracle)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_bipartite(b.row_off, b.nbrs)
    for L, _ in b.cases:
        want = {k: b.z[f"{k}_L{L}"] for k in ("ids", "dists", "cmps", "hops")}
        assert_same_search(idx.search_bipartite(queries, b.k, L, init_ids=b.z[f"init_ids_L{L}"]), want, f"{golden.name} L={L}")
    idx.close()


@pytest.mark.skipif(not have_detrand_reference(), reason="oracle/_ref/libref_roargraph_detrand.so not available")
def test_bipartite_parity_vs_compiled_reference(golden, oracle, tmp_path):
    """A2 pinned by execution: the reference's own SearchBipartiteGraph (compiled with a fixed std::random_device word per
    query, oracle/shims/fixed_random_device.h) and the GPU kernel given the entry points that word yields: ids, distance
    bits, cmps, hops."""
    from oracle.oracle import Reference
    from roargraph_b200 import formats as F
    R = Reference(detrand=True)
    base, queries = golden.normalised(oracle)
    nd = base.shape[0]
    row_off, nbrs = bipartite_from_knn(golden.z["train_knn_ids"], nd, 16)
    bf, gf = str(tmp_path / "base.fbin"), str(tmp_path / "bp.rg")
    F.write_fbin(bf, golden.base_raw)
    F.write_bipartite_index(gf, row_off, nbrs)
    h = R.bipartite_load(bf, gf, golden.metric, base.shape[1], golden.z["train_knn_ids"].shape[0], 2)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_bipartite(row_off, nbrs)
    checked = 0
    for L, seed in ((10, 1), (50, 2), (300, 3), (2000, 4)):
        try:
            want = R.search_bipartite(h, queries, golden.k, L, 2, seed_base=seed)
        except RuntimeError as e:  # some query's entry points reach fewer than k nodes: both sides must fail alike
            assert "not enough results" in str(e)
            init = np.stack([R.bipartite_entry_points(seed + i, nd) for i in range(queries.shape[0])])
            with pytest.raises(rg.RgError) as ge:
                idx.search_bipartite(queries, golden.k, L, init_ids=init)
            assert ge.value.code == rg.RG_ENOTENOUGH
            continue
        assert_same_search(idx.search_bipartite(queries, golden.k, L, init_ids=want["init_ids"]), want,
                           f"bipartite vs compiled reference L={L}")
        checked += 1
    assert checked >= 2
    R.index_free(h)
    idx.close()


@pytest.mark.skipif(not have_detrand_reference(), reason="oracle/_ref/libref_roargraph_detrand.so not available")
@pytest.mark.parametrize("nb,dim,metric", [(60000, 200, "ip"), (30000, 96, "l2")])
def test_bipartite_live_reference_parity_at_scale(nb, dim, metric, tmp_path):
    """The same pin on a graph of README shape (M_pjbp = 35, as many training queries as base rows so that most base
    nodes own a query node): 500 queries over the sweep's range, every id / distance bit / cmps / hops."""
    from oracle.oracle import Reference
    from roargraph_b200 import formats as F, synth
    R = Reference(detrand=True)
    base, train, test = synth.make_numpy(nb, nb, 500, dim, seed=77)
    knn, _ = rg.knn(base, train, 35, metric)
    row_off, nbrs = bipartite_from_knn(knn.astype(np.uint32), nb, 35)
    bf, gf = str(tmp_path / "base.fbin"), str(tmp_path / "bp.rg")
    F.write_fbin(bf, base)
    F.write_bipartite_index(gf, row_off, nbrs)
    threads = R.num_procs()
    h = R.bipartite_load(bf, gf, metric, F.pad_rows(base).shape[1], nb, threads)
    b = R.index_base(h, nb, F.pad_rows(base).shape[1]).copy()
    q = F.pad_rows(test)
    idx = rg.RoarIndex(b, metric)
    idx.set_bipartite(row_off, nbrs)
    for L in (10, 50, 200, 1000):
        want = R.search_bipartite(h, q, 10, L, threads, seed_base=1000 * L)
        assert_same_search(idx.search_bipartite(q, 10, L, init_ids=want["init_ids"]), want, f"{metric} nb={nb} L={L}")
    R.index_free(h)
    idx.close()


def test_bipartite_not_enough_results(oracle, golden):
    """Ten entry points on one base node without edges: one result, the reference's exception text (:345-349)."""
    base, queries = golden.normalised(oracle)
    nd = base.shape[0]
    row_off, nbrs = bipartite_from_knn(golden.z["train_knn_ids"], nd, 12)
    lonely = int(np.flatnonzero(np.diff(row_off.astype(np.int64))[:nd] == 0)[0])
    init = np.full((queries.shape[0], 10), lonely, dtype=np.uint32)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_bipartite(row_off, nbrs)
    with pytest.raises(rg.RgError) as e:
        idx.search_bipartite(queries, 2, 20, init_ids=init)
    assert e.value.code == rg.RG_ENOTENOUGH and "not enough results: 1, expected: 2" in str(e.value)
    ids, dists, cmps, hops = idx.search_bipartite(queries, 1, 20, init_ids=init)
    assert (ids == lonely).all() and (cmps == 0).all() and (hops == 1).all()
    idx.close()


@pytest.mark.parametrize("R", ["8", "16"])
def test_merge_insert_equals_one_by_one_insert(oracle, golden, monkeypatch, R):
    """RG_QUEUE_MERGE=0 (one shift per candidate) and =1 (one merge pass per row group) are observably identical, and
    both equal the oracle — short queues that overflow inside a group, long ones that never fill."""
    base, queries = golden.normalised(oracle)
    monkeypatch.setenv("RG_ROWS_PER_BATCH", R)
    for L in (10, 13, 40, 700, 2000):
        want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries, golden.k, L)
        for merge in ("0", "1"):
            monkeypatch.setenv("RG_QUEUE_MERGE", merge)
            idx = rg.RoarIndex(base, golden.metric)
            idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
            assert_same_search(idx.search(queries, golden.k, L), want, f"merge={merge} R={R} L={L}")
            idx.close()


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not available")
@pytest.mark.parametrize("nb,dim,metric", [(60000, 200, "ip"), (30000, 512, "cosine"), (30000, 96, "l2")])
def test_live_reference_parity_at_scale(nb, dim, metric):
    """Index built by the reference's own BuildRoarGraph on the box; GPU search vs the reference's SearchRoarGraph
    (compiled unmodified) on 500 queries across the L_pq sweep's range."""
    c = build_reference_case(nb, nb // 2, 500, dim, metric)
    idx = rg.RoarIndex(c["base"], metric)
    idx.set_graph(c["ep"], c["row_off"], c["nbrs"])
    for L in (10, 20, 50, 100, 200, 500, 1000, 2000):
        want = c["R"].search_roargraph(c["h"], c["queries"], 10, L, c["threads"])
        assert_same_search(idx.search(c["queries"], 10, L), want, f"{metric} nb={nb} L={L}")
    c["R"].index_free(c["h"])
    idx.close()


def test_index_from_fbin_streams_pads_and_hooks(tmp_path, oracle, golden):
    """rg_index_create_from_fbin: the file streamed through the pinned staging buffers gives the index the padded (and, with
    the row hook, normalised) rows a host-side load would have given it — same ids, distance bits, cmps, hops."""
    from roargraph_b200 import formats as F
    base, queries = golden.normalised(oracle)
    bf = str(tmp_path / "base.fbin")
    F.write_fbin(bf, golden.base_raw)
    hook = None
    if golden.metric == "cosine":
        raw_dim = golden.base_raw.shape[1]
        want_rows = base[:, :raw_dim]

        def hook(rows):  # stand-in for the mirror's normalize_rows: hand the library the rows the reference's loader produces
            lo = hook.done
            rows[:] = want_rows[lo:lo + rows.shape[0]]
            hook.done += rows.shape[0]
        hook.done = 0
    idx = rg.RoarIndex.from_fbin(bf, golden.metric, row_hook=hook)
    assert (idx.nd, idx.dim) == base.shape
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    for L in golden.Ls:
        want = {k: golden.z[f"{k}_L{L}"] for k in ("ids", "dists", "cmps", "hops")}
        assert_same_search(idx.search(queries, golden.k, L), want, f"from_fbin {golden.name} L={L}")
    idx.close()
    # a dimension that is not a multiple of 8: padded with zeros by the copy, distances unchanged
    rng = np.random.default_rng(4)
    b13 = rng.standard_normal((3000, 13)).astype(np.float32)
    q13 = F.pad_rows(rng.standard_normal((20, 13)).astype(np.float32))
    F.write_fbin(bf, b13)
    row_off, nbrs = random_graph(rng, 3000, 4, 12)
    a = rg.RoarIndex.from_fbin(bf, "l2")
    b = rg.RoarIndex(F.pad_rows(b13), "l2")
    assert a.dim == 16
    for x in (a, b):
        x.set_graph(1, row_off, nbrs)
    ra, rb = a.search(q13, 5, 30), b.search(q13, 5, 30)
    assert all((u == v).all() for u, v in zip(ra[:1] + ra[2:], rb[:1] + rb[2:])) and (bits(ra[1]) == bits(rb[1])).all()
    a.close()
    b.close()
    with pytest.raises(rg.RgError):
        rg.RoarIndex.from_fbin(str(tmp_path / "missing.fbin"), "l2")
    open(bf, "ab").write(b"xx")  # trailing bytes: "Data file size wrong!" like util.h:124
    with pytest.raises(rg.RgError) as e:
        rg.RoarIndex.from_fbin(bf, "l2")
    assert "Data file size wrong" in str(e.value)


def test_bipartite_graph_must_be_bipartite(oracle, golden):
    """base -> base or query -> query edges would make the kernel read past the base array / the visited tags: rejected."""
    base, _ = golden.normalised(oracle)
    nd = base.shape[0]
    idx = rg.RoarIndex(base, golden.metric)
    row_off = np.zeros(nd + 3, dtype=np.uint64)
    for bad in ([5], [nd + 1]):            # base node 0 -> base node 5; query node nd -> query node nd + 1
        ro = row_off.copy()
        if bad[0] < nd:
            ro[1:] = 1
        else:
            ro[nd + 1:] = 1
        with pytest.raises(rg.RgError) as e:
            idx.set_bipartite(ro, np.array(bad, dtype=np.uint32))
        assert e.value.code == rg.RG_EINVAL
    idx.close()


@pytest.mark.parametrize("min_l", ["0", "10"])
def test_lazy_queue_equals_eager_queue(oracle, golden, monkeypatch, min_l):
    """RG_LAZY_QUEUE_MIN_L: the pending-block queue (default for L_pq >= 256, forced from L_pq = 10 here) and the eager
    sorted-array queue (0 = never lazy) are observably identical — and both equal the oracle — for short queues that overflow
    inside a flush, long ones that never fill, k == L, and the bipartite search."""
    base, queries = golden.normalised(oracle)
    monkeypatch.setenv("RG_LAZY_QUEUE_MIN_L", min_l)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    for L in (10, 13, 31, 40, 64, 257, 700, 2000):
        want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries, golden.k, L)
        assert_same_search(idx.search(queries, golden.k, L), want, f"min_l={min_l} L={L}")
    want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries[:9], 40, 40)
    if not want["nfail"]:
        assert_same_search(idx.search(queries[:9], 40, 40), want, "k == L")
    row_off, nbrs = bipartite_from_knn(golden.z["train_knn_ids"], base.shape[0], 12)
    init = np.random.default_rng(3).integers(0, base.shape[0], (queries.shape[0], 10)).astype(np.uint32)
    idx.set_bipartite(row_off, nbrs)
    for L in (10, 50, 300):
        want = oracle.search_bipartite(base, row_off, nbrs, init, golden.metric, queries, 5, L)
        if not want["nfail"]:
            assert_same_search(idx.search_bipartite(queries, 5, L, init_ids=init), want, f"bipartite min_l={min_l} L={L}")
    idx.close()


def test_query_ordering_is_unobservable(oracle, golden, monkeypatch):
    """RG_QUERY_ORDER=1 hands batches of >= 4 096 queries to the warps grouped by their nearest landmark row (an opt-in
    experiment: no gain measured); every query is independent, so ids, distance bits, cmps and hops equal the arrival-order
    run and the oracle, query by query."""
    base, queries = golden.normalised(oracle)
    reps = -(-5000 // queries.shape[0])
    big = np.ascontiguousarray(np.tile(queries, (reps, 1))[:5000])
    # make the rows distinct so that the ordering has something to sort (tiny perturbation, still the same graph)
    big = (big + np.float32(1e-3) * np.random.default_rng(1).standard_normal(big.shape).astype(np.float32)).astype(np.float32)
    if golden.metric == "cosine":
        big /= np.linalg.norm(big, axis=1, keepdims=True)
    idx = rg.RoarIndex(base, golden.metric)
    idx.set_graph(golden.ep, golden.row_off, golden.nbrs)
    L = 40
    want = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, big, golden.k, L)
    monkeypatch.setenv("RG_QUERY_ORDER", "1")
    got_ordered = idx.search(big, golden.k, L)
    monkeypatch.setenv("RG_QUERY_ORDER", "0")
    idx2 = rg.RoarIndex(base, golden.metric)
    idx2.set_graph(golden.ep, golden.row_off, golden.nbrs)
    got_plain = idx2.search(big, golden.k, L)
    assert_same_search(got_ordered, want, "ordered tickets")
    assert_same_search(got_plain, want, "arrival order")
    idx.close()
    idx2.close()
