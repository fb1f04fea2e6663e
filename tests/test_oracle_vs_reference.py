"""CPU, build container only: the oracle restatement against the reference library on fresh random inputs
(skipped where oracle/_ref has not been built)."""
import numpy as np
import pytest

from oracle.oracle import have_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built")


def test_distance_random(oracle):
    from oracle.oracle import Reference
    R = Reference()
    rng = np.random.default_rng(3)
    for dim in (8, 16, 40, 100, 203, 512):  # includes dims with 4-block / masked tails (unused on the search path)
        for _ in range(200):
            a = rng.standard_normal(dim).astype(np.float32)
            b = rng.standard_normal(dim).astype(np.float32)
            for m in ("ip", "l2"):
                assert np.float32(R.distance(m, a, b)).view(np.uint32) == np.float32(oracle.distance(m, a, b)).view(np.uint32)


def test_normalize_matches(oracle):
    from oracle.oracle import Reference
    R = Reference()
    rng = np.random.default_rng(5)
    for dim in (8, 64, 104, 200, 512):
        x = (rng.standard_normal((300, dim)) * rng.uniform(0.01, 100.0, (300, 1))).astype(np.float32)
        # the restatement follows the COMPILED loop (lane sums, RSQRTSS + Newton step, multiply): bit for bit
        assert (R.normalize_rows(x.copy()).view(np.uint32) == oracle.normalize_rows(x.copy()).view(np.uint32)).all(), dim
    x = rng.standard_normal((50, 13)).astype(np.float32)  # dim % 8 != 0: outside the reference's precondition, rounding only
    np.testing.assert_allclose(R.normalize_rows(x.copy()), oracle.normalize_rows(x.copy()), rtol=2e-6, atol=1e-7)


def test_search_live_reference(oracle, golden, tmp_path):
    """Load the golden index through the reference's own loaders and compare live (not stored) results."""
    from oracle.oracle import Reference
    from roargraph_b200 import formats as F
    R = Reference()
    bf, gf = str(tmp_path / "base.fbin"), str(tmp_path / "index.rg")
    F.write_fbin(bf, golden.base_raw)
    golden.z["index_words"].tofile(gf)
    h = R.index_load(bf, gf, golden.metric, golden.base.shape[1], 2)
    base, queries = golden.normalised(oracle)
    assert (R.index_base(h, base.shape[0], base.shape[1]).view(np.uint32) == base.view(np.uint32)).all()
    for L in (golden.k, 17, 64):
        r = R.search_roargraph(h, queries, golden.k, L, 2)
        o = oracle.search_roargraph(base, golden.row_off, golden.nbrs, golden.ep, golden.metric, queries, golden.k, L, 2)
        assert (r["ids"] == o["ids"]).all() and (r["dists"].view(np.uint32) == o["dists"].view(np.uint32)).all()
        assert (r["cmps"] == o["cmps"]).all() and (r["hops"] == o["hops"]).all()
    R.index_free(h)


# ---- SearchBipartiteGraph (src/index_bipartite.cpp:282-356): executed pin -----------------------------------------------
def _bipartite_case(golden, tmp_path, M_pjbp=16):
    from gpu_helpers import bipartite_from_knn
    from roargraph_b200 import formats as F
    nd = golden.base_raw.shape[0]
    knn = golden.z["train_knn_ids"]
    row_off, nbrs = bipartite_from_knn(knn, nd, M_pjbp)
    bf, gf = str(tmp_path / "base.fbin"), str(tmp_path / "bipartite.rg")
    F.write_fbin(bf, golden.base_raw)
    F.write_bipartite_index(gf, row_off, nbrs)
    return bf, gf, row_off, nbrs, knn.shape[0]


@pytest.mark.skipif(not __import__("oracle.oracle", fromlist=["x"]).have_detrand_reference(),
                    reason="oracle/_ref/libref_roargraph_detrand.so not built")
def test_search_bipartite_restatement_equals_the_compiled_reference(oracle, golden, tmp_path):
    """The compiled reference with a fixed std::random_device word per query (oracle/shims/fixed_random_device.h) against
    the plain-C restatement given the SAME 10 entry points: ids, distance bits, cmps and hops of every query."""
    from oracle.oracle import Reference
    R = Reference(detrand=True)
    bf, gf, row_off, nbrs, nsq = _bipartite_case(golden, tmp_path)
    base, queries = golden.normalised(oracle)
    nd = base.shape[0]
    h = R.bipartite_load(bf, gf, golden.metric, base.shape[1], nsq, 2)
    assert (R.index_base(h, nd, base.shape[1]).view(np.uint32) == base.view(np.uint32)).all()
    checked = 0
    for L, seed in ((10, 0), (33, 1000), (64, 77), (200, 123456)):
        try:
            r = R.search_bipartite(h, queries, golden.k, L, 2, seed_base=seed)
        except RuntimeError as e:  # a query whose 10 entry points reach < k nodes: the reference throws (:346-350)
            assert "not enough results" in str(e)
            continue
        # the entry points the harness reports are what libstdc++ derives from the seed word, and they lie in [0, nd)
        assert (r["init_ids"][3] == R.bipartite_entry_points(seed + 3, nd)).all() and (r["init_ids"] < nd).all()
        o = oracle.search_bipartite(base, row_off, nbrs, r["init_ids"], golden.metric, queries, golden.k, L, 2)
        assert o["nfail"] == 0
        assert (r["ids"] == o["ids"]).all(), f"L={L}"
        assert (r["dists"].view(np.uint32) == o["dists"].view(np.uint32)).all(), f"L={L}"
        assert (r["cmps"] == o["cmps"]).all() and (r["hops"] == o["hops"]).all(), f"L={L}"
        checked += 1
    assert checked >= 2
    # the same seed word reproduces the same answer; another word changes the entry points
    a = R.search_bipartite(h, queries[:16], golden.k, 64, 2, seed_base=5)
    b = R.search_bipartite(h, queries[:16], golden.k, 64, 1, seed_base=5)
    c = R.search_bipartite(h, queries[:16], golden.k, 64, 2, seed_base=6)
    assert (a["ids"] == b["ids"]).all() and (a["cmps"] == b["cmps"]).all()
    assert (a["init_ids"] != c["init_ids"]).any()
    R.index_free(h)
