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
