"""CPU: file formats of the reference round-trip and decode the reference-written index bytes."""
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
