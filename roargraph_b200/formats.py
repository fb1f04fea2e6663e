"""On-disk formats of the reference, read/written with numpy (host-side plumbing).

* fbin  : ``u32 n, u32 d, n*d f32``                         (/root/reference/README.md:14, include/efanna2e/util.h:107-211)
* projection index : ``u32 ep, u32 npts, {u32 deg, u32 ids[deg]} * npts``
                                                            (src/index_bipartite.cpp:2097-2117 Load, :2606-2619 Save)
* bipartite index  : ``u32 npts, {u32 deg, u32 ids[deg]} * npts``   (src/index_bipartite.cpp:2045-2071)
* truthset / learn-base kNN : ``i32 n, i32 K, n*K u32 ids, n*K f32 dists``
                                                            (thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:325-343;
                                                             readers include/efanna2e/util.h:85-105,130-155 and
                                                             src/index_bipartite.cpp:2622-2642 which reads ids only)
"""
from __future__ import annotations

import numpy as np


def write_fbin(path: str, x: np.ndarray) -> None:
    x = np.ascontiguousarray(x, dtype=np.float32)
    with open(path, "wb") as f:
        np.array([x.shape[0], x.shape[1]], dtype=np.uint32).tofile(f)
        x.tofile(f)


def read_fbin(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        n, d = np.fromfile(f, dtype=np.uint32, count=2)
        x = np.fromfile(f, dtype=np.float32, count=int(n) * int(d))
    if x.size != int(n) * int(d):
        raise RuntimeError("Data file size wrong!")  # util.h:124
    return x.reshape(int(n), int(d))


def pad_rows(x: np.ndarray, factor: int = 8) -> np.ndarray:
    """Row stride = dim rounded up to a multiple of 8 with zero fill (util.h:37-75 data_align, :191-199)."""
    n, d = x.shape
    nd = (d + factor - 1) // factor * factor
    if nd == d:
        return np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((n, nd), dtype=np.float32)
    out[:, :d] = x
    return out


def _graph_to_csr(buf: np.ndarray, npts: int):
    """Decode a ``{deg, ids[deg]}*npts`` u32 stream into (row_off u64[npts+1], nbrs u32[])."""
    row_off = np.zeros(npts + 1, dtype=np.uint64)
    pos = 0
    # vectorised decode is impossible without the degrees; walk once (npts is at most 10M -> use a small C-like loop
    # through numpy views in chunks).
    degs = np.empty(npts, dtype=np.uint32)
    starts = np.empty(npts, dtype=np.int64)
    for i in range(npts):
        d = int(buf[pos])
        degs[i] = d
        starts[i] = pos + 1
        pos += 1 + d
    row_off[1:] = np.cumsum(degs, dtype=np.uint64)
    total = int(row_off[-1])
    nbrs = np.empty(total, dtype=np.uint32)
    # gather: positions of neighbour words = all positions except the degree slots
    mask = np.ones(pos, dtype=bool)
    mask[starts - 1] = False
    nbrs[:] = buf[:pos][mask]
    return row_off, nbrs


def read_projection_index(path: str):
    """-> (ep, row_off u64[npts+1], nbrs u32[total])   src/index_bipartite.cpp:2097-2117"""
    buf = np.fromfile(path, dtype=np.uint32)
    ep, npts = int(buf[0]), int(buf[1])
    row_off, nbrs = _graph_to_csr(buf[2:], npts)
    return ep, row_off, nbrs


def write_projection_index(path: str, ep: int, row_off: np.ndarray, nbrs: np.ndarray) -> None:
    """src/index_bipartite.cpp:2606-2619"""
    npts = len(row_off) - 1
    degs = np.diff(row_off.astype(np.int64)).astype(np.uint32)
    out = np.empty(2 + npts + len(nbrs), dtype=np.uint32)
    out[0], out[1] = ep, npts
    pos = (np.arange(npts, dtype=np.int64) + row_off[:-1].astype(np.int64)) + 2  # slot of each degree word
    out[pos] = degs
    mask = np.ones(len(out), dtype=bool)
    mask[:2] = False
    mask[pos] = False
    out[mask] = nbrs
    out.tofile(path)


def read_bipartite_index(path: str):
    """-> (row_off, nbrs) over nd + nd_sq nodes   src/index_bipartite.cpp:2058-2071"""
    buf = np.fromfile(path, dtype=np.uint32)
    npts = int(buf[0])
    return _graph_to_csr(buf[1:], npts)


def write_bipartite_index(path: str, row_off: np.ndarray, nbrs: np.ndarray) -> None:
    npts = len(row_off) - 1
    degs = np.diff(row_off.astype(np.int64)).astype(np.uint32)
    out = np.empty(1 + npts + len(nbrs), dtype=np.uint32)
    out[0] = npts
    pos = (np.arange(npts, dtype=np.int64) + row_off[:-1].astype(np.int64)) + 1
    out[pos] = degs
    mask = np.ones(len(out), dtype=bool)
    mask[0] = False
    mask[pos] = False
    out[mask] = nbrs
    out.tofile(path)


def write_truthset(path: str, ids: np.ndarray, dists: np.ndarray) -> None:
    """compute_groundtruth.cpp:325-343 (ids then distances, both n*K)."""
    n, k = ids.shape
    with open(path, "wb") as f:
        np.array([n, k], dtype=np.int32).tofile(f)
        np.ascontiguousarray(ids, dtype=np.uint32).tofile(f)
        np.ascontiguousarray(dists, dtype=np.float32).tofile(f)


def read_truthset(path: str):
    with open(path, "rb") as f:
        n, k = np.fromfile(f, dtype=np.int32, count=2)
        ids = np.fromfile(f, dtype=np.uint32, count=int(n) * int(k)).reshape(int(n), int(k))
        d = np.fromfile(f, dtype=np.float32, count=int(n) * int(k))
    dists = d.reshape(int(n), int(k)) if d.size == int(n) * int(k) else None
    return ids, dists


def compute_recall(res: np.ndarray, gt: np.ndarray, k: int) -> float:
    """tests/test_search_roargraph.cpp:23-36 ComputeRecall: |gt[:k] ∩ res[:k]| / (k * nq)."""
    hit = 0
    for r, g in zip(res[:, :k], gt[:, :k]):
        hit += np.isin(g, r).sum()
    return float(hit) / float(k * res.shape[0])
