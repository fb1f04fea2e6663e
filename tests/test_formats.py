"""CPU: file formats of the reference round-trip and decode the reference-written index bytes."""
import os

import numpy as np

from conftest import GoldenCase
from roargraph_b200 import formats as F


def test_projection_index_roundtrip(tmp_path):
    g = GoldenCase("ip_d40")
    p = str(tmp_path / "index.rg")
    g.z["index_words"].tofile(p)
    ep, ro, nb = F.read_projection_index(p)
    assert ep == g.ep and (ro == g.row_off).all() and (nb == g.nbrs).all()
    p2 = str(tmp_path / "index2.rg")
    F.write_projection_index(p2, ep, ro, nb)
    assert (np.fromfile(p2, dtype=np.uint32) == g.z["index_words"]).all()


def test_fbin_truthset_bipartite_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((17, 13)).astype(np.float32)
    F.write_fbin(str(tmp_path / "x.fbin"), x)
    assert (F.read_fbin(str(tmp_path / "x.fbin")) == x).all()
    assert F.pad_rows(x).shape == (17, 16) and (F.pad_rows(x)[:, 13:] == 0).all()
    ids = rng.integers(0, 100, (5, 7)).astype(np.uint32)
    d = rng.standard_normal((5, 7)).astype(np.float32)
    F.write_truthset(str(tmp_path / "gt"), ids, d)
    i2, d2 = F.read_truthset(str(tmp_path / "gt"))
    assert (i2 == ids).all() and (d2 == d).all()
    ro = np.array([0, 2, 2, 5], dtype=np.uint64)
    nb = np.array([1, 2, 0, 1, 2], dtype=np.uint32)
    F.write_bipartite_index(str(tmp_path / "bp"), ro, nb)
    ro2, nb2 = F.read_bipartite_index(str(tmp_path / "bp"))
    assert (ro2 == ro).all() and (nb2 == nb).all()


def test_recall():
    res = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.uint32)
    gt = np.array([[3, 2, 9, 1], [7, 8, 9, 4]], dtype=np.uint32)
    assert abs(F.compute_recall(res, gt, 3) - 2 / 6) < 1e-9


def test_ragged_graph_files_roundtrip_property(tmp_path):
    """Random ragged adjacency (empty rows, an all-empty graph, one long row) through both index file formats."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=0, max_value=9), min_size=1, max_size=40), st.integers(0, 2**31 - 1))
    def check(degs, seed):
        rng = np.random.default_rng(seed)
        degs = np.array(degs, dtype=np.int64)
        if seed % 5 == 0:
            degs[seed % len(degs)] = 300  # one long row
        ro = np.zeros(len(degs) + 1, dtype=np.uint64)
        ro[1:] = np.cumsum(degs)
        nb = rng.integers(0, len(degs), int(ro[-1])).astype(np.uint32)
        ep = int(seed % len(degs))
        p = str(tmp_path / "g.rg")
        F.write_projection_index(p, ep, ro, nb)
        assert os.path.getsize(p) == 4 * (2 + len(degs) + len(nb))
        ep2, ro2, nb2 = F.read_projection_index(p)
        assert ep2 == ep and (ro2 == ro).all() and (nb2 == nb).all()
        F.write_bipartite_index(p, ro, nb)
        ro3, nb3 = F.read_bipartite_index(p)
        assert (ro3 == ro).all() and (nb3 == nb).all()

    check()
