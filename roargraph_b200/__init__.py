"""roargraph_b200 — B200-native (sm_100a) implementation of RoarGraph's distance-bound hot path.

This package is a thin ctypes binding over the C-ABI shared library ``librg_b200.so`` declared in
``include/rg_b200.h`` (hand-written CUDA kernels, no CPU fallback).  It is what tests/ and bench.py call; the C++
host mirror of the reference's ``efanna2e::IndexBipartite`` surface and the re-hosted CLIs live in ``host/``.

Reference paths replaced (relative to /root/reference):
  * IndexBipartite::SearchRoarGraph        src/index_bipartite.cpp:2311-2420
  * IndexBipartite::SearchBipartiteGraph   src/index_bipartite.cpp:282-356
  * exact_knn                              thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:126-248
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librg_b200.so")

METRIC_L2, METRIC_INNER_PRODUCT, METRIC_COSINE = 0, 1, 4  # efanna2e::Metric (include/efanna2e/distance.h:15)
METRIC = {"l2": METRIC_L2, "ip": METRIC_INNER_PRODUCT, "cosine": METRIC_COSINE,
          "mips": METRIC_INNER_PRODUCT}

RG_OK, RG_EINVAL, RG_ECUDA, RG_ENOTENOUGH, RG_ESTATE, RG_ENOMEM = 0, -1, -2, -3, -4, -5


class RgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile librg_b200.so in-tree with nvcc for sm_100a (roargraph_b200/csrc/Makefile)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j", "4"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library.  Fails loudly when it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RgError(RG_ECUDA, f"{LIB_PATH} is missing: run roargraph_b200.build() / `make -C roargraph_b200/csrc` "
                                    "(the product path has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
        L.rg_last_error.restype = C.c_char_p
        L.rg_abi_version.restype = C.c_int
        L.rg_init.argtypes = [C.c_int]
        L.rg_device_sm_count.restype = C.c_int
        L.rg_index_create.argtypes = [C.POINTER(vp), vp, u64, u32, C.c_int, C.c_int]
        L.rg_index_create_from_fbin.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int, vp, vp, C.POINTER(u64), C.POINTER(u32)]
        L.rg_index_set_graph.argtypes = [vp, u32, vp, vp]
        L.rg_index_set_bipartite.argtypes = [vp, u64, vp, vp]
        L.rg_index_destroy.argtypes = [vp]
        L.rg_index_build_roargraph.argtypes = [vp, vp, C.c_int, u64, u32, u32, u32, u32, u32]
        L.rg_index_graph_size.argtypes = [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(u32)]
        L.rg_index_get_graph.argtypes = [vp, vp, vp]
        L.rg_search_roargraph.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp, vp]
        L.rg_search_roargraph_dev.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp, vp, vp]
        L.rg_index_check.argtypes = [vp, vp]
        L.rg_search_bipartite.argtypes = [vp, vp, u32, u32, u32, vp, u32, u64, vp, vp, vp, vp]
        L.rg_search_bipartite_dev.argtypes = [vp, vp, u32, u32, u32, vp, u32, u64, vp, vp, vp, vp, vp]
        L.rg_search_last_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.rg_search_last_launch.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.rg_knn.argtypes = [vp, u64, vp, u64, u32, u32, C.c_int, i32, vp, vp]
        L.rg_knn_dev.argtypes = [vp, u64, vp, u64, u32, u32, C.c_int, i32, vp, vp, vp]
        L.rg_topk_merge_dev.argtypes = [vp, vp, u32, u64, u32, vp, vp, vp]
        L.rg_topk_pack_dev.argtypes = [vp, vp, u64, u32, vp, vp]
        L.rg_topk_merge_packed_dev.argtypes = [vp, u32, u64, u32, u32, u64, u64, vp, vp, vp]
        L.rg_group_create.argtypes = [C.POINTER(vp), C.c_int, vp]
        L.rg_group_destroy.argtypes = [vp]
        L.rg_group_size.argtypes = [vp]
        L.rg_group_knn.argtypes = [vp, vp, u64, vp, u64, u32, u32, C.c_int, vp, vp]
        L.rg_sharded_index_create.argtypes = [vp, C.POINTER(vp), vp, u64, u32, C.c_int]
        L.rg_sharded_index_destroy.argtypes = [vp]
        L.rg_sharded_index_shards.argtypes = [vp]
        L.rg_sharded_index_shard.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]
        L.rg_sharded_index_build_roargraph.argtypes = [vp, vp, u64, u32, u32, u32]
        L.rg_sharded_search_roargraph.argtypes = [vp, vp, u32, u32, u32, vp, vp]
        L.rg_knn_base_create.argtypes = [C.POINTER(vp), vp, u64, u32, C.c_int, C.c_int]
        L.rg_knn_base_destroy.argtypes = [vp]
        L.rg_knn_base_search_dev.argtypes = [vp, vp, u64, u32, i32, vp, vp, vp]
        L.rg_knn_base_seed_dev.argtypes = [vp, vp, u64, u32, u32, vp, vp]
        L.rg_knn_base_search_seeded_dev.argtypes = [vp, vp, u64, u32, i32, vp, u64, vp, vp, vp]
        L.rg_knn_base_search_resync_dev.argtypes = [vp, vp, u64, u32, i32, vp, u64, u32, u64, vp, KNN_BOUND_HOOK, vp, vp, vp, vp]
        L.rg_knn_profile.argtypes = [C.POINTER(C.c_double)]
        _lib = L
    return _lib


# int hook(void *user, uint64_t q_off, uint64_t n, void *stream)  (include/rg_b200.h: rg_knn_bound_hook)
KNN_BOUND_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p)


def _check(rc: int) -> None:
    if rc != 0:
        raise RgError(rc, lib().rg_last_error().decode())


def _np_ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def init(device: int = 0) -> None:
    _check(lib().rg_init(device))


def sm_count() -> int:
    n = lib().rg_device_sm_count()
    if n < 0:
        _check(n)
    return n


def _stream_ptr(stream) -> C.c_void_p:
    if stream is None:
        return C.c_void_p(0)
    return C.c_void_p(int(getattr(stream, "cuda_stream", stream)))


class RoarIndex:
    """Device-resident base vectors + graph(s); mirrors IndexBipartite's search-side surface for a whole batch.

    ``base`` is either a C-contiguous float32 numpy array (copied to HBM once) or a CUDA torch tensor (used in
    place; must outlive the index).  Row length must already be padded to a multiple of 8 (formats.pad_rows), and
    rows must already be normalised for ``cosine`` — exactly what the reference's loader hands to its search
    (src/index_bipartite.cpp:2664-2695).
    """

    def __init__(self, base, metric: str | int):
        self._h = C.c_void_p(0)
        m = METRIC[metric] if isinstance(metric, str) else int(metric)
        self.metric = m
        if isinstance(base, np.ndarray):
            if base.dtype != np.float32 or not base.flags.c_contiguous or base.ndim != 2:
                raise RgError(RG_EINVAL, "base must be a C-contiguous float32 matrix")
            self._keep = None
            self.nd, self.dim = base.shape
            _check(lib().rg_index_create(C.byref(self._h), _np_ptr(base), self.nd, self.dim, m, 0))
        else:  # torch CUDA tensor
            if not base.is_cuda or not base.is_contiguous() or base.dim() != 2 or str(base.dtype) != "torch.float32":
                raise RgError(RG_EINVAL, "base must be a contiguous float32 CUDA tensor")
            self._keep = base
            self.nd, self.dim = base.shape
            _check(lib().rg_index_create(C.byref(self._h), C.c_void_p(base.data_ptr()), self.nd, self.dim, m, 1))

    ROW_HOOK = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_uint64, C.c_uint32, C.c_void_p)

    @classmethod
    def from_fbin(cls, path: str, metric: str | int, row_hook=None) -> "RoarIndex":
        """The base streamed from an fbin file into HBM (rg_index_create_from_fbin: pinned double buffer, rows padded to x8
        by the copy engine).  row_hook(rows: float32 [n, dim] numpy view of a staged slab, modified in place) is applied
        to every slab before it is copied — e.g. the reference's row normalisation for cosine."""
        self = cls.__new__(cls)
        self._h, self._keep, self._owned = C.c_void_p(0), None, True
        m = METRIC[metric] if isinstance(metric, str) else int(metric)
        self.metric = m
        cb = None
        if row_hook is not None:
            def _cb(ptr, n, dim, _user):
                row_hook(np.ctypeslib.as_array(ptr, shape=(int(n), int(dim))))
            cb = cls.ROW_HOOK(_cb)
        nd, dim = C.c_uint64(0), C.c_uint32(0)
        _check(lib().rg_index_create_from_fbin(C.byref(self._h), path.encode(), m, C.cast(cb, C.c_void_p) if cb else None, None,
                                               C.byref(nd), C.byref(dim)))
        self.nd, self.dim = nd.value, dim.value
        return self

    @classmethod
    def _borrowed(cls, handle, nd: int, dim: int, metric: int) -> "RoarIndex":
        """A view of an rg_index owned by someone else (a shard of a GroupShardedIndex): close() does not destroy it."""
        self = cls.__new__(cls)
        self._h, self._keep, self._owned = C.c_void_p(handle), None, False
        self.nd, self.dim, self.metric = nd, dim, metric
        return self

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            if getattr(self, "_owned", True):
                lib().rg_index_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- graphs ------------------------------------------------------------------------------------------------
    def set_graph(self, ep: int, row_off: np.ndarray, nbrs: np.ndarray) -> None:
        """LoadProjectionGraph's content as CSR (src/index_bipartite.cpp:2097-2117)."""
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64)
        nbrs = np.ascontiguousarray(nbrs, dtype=np.uint32)
        if row_off.shape[0] != self.nd + 1:
            raise RgError(RG_EINVAL, "row_off must have nd + 1 entries")
        _check(lib().rg_index_set_graph(self._h, int(ep), _np_ptr(row_off), _np_ptr(nbrs)))

    def set_bipartite(self, row_off: np.ndarray, nbrs: np.ndarray) -> None:
        """IndexBipartite::Load's content as CSR over nd + nd_sq nodes (src/index_bipartite.cpp:2058-2071)."""
        row_off = np.ascontiguousarray(row_off, dtype=np.uint64)
        nbrs = np.ascontiguousarray(nbrs, dtype=np.uint32)
        _check(lib().rg_index_set_bipartite(self._h, row_off.shape[0] - 1, _np_ptr(row_off), _np_ptr(nbrs)))

    def build_roargraph(self, knn_ids, M_sq: int = 100, M_pjbp: int = 35, L_pjpq: int = 500, n_rounds: int = 0) -> None:
        """BuildRoarGraph on the device (src/index_bipartite.cpp:143-233) from the training queries' kNN id lists
        (numpy uint32/int32 [nsq, K] or a CUDA int32 tensor); parameters as tests/test_build_roargraph.cpp:52-60."""
        if isinstance(knn_ids, np.ndarray):
            a = np.ascontiguousarray(knn_ids).view(np.uint32)
            _check(lib().rg_index_build_roargraph(self._h, _np_ptr(a), 0, a.shape[0], a.shape[1], M_sq, M_pjbp, L_pjpq,
                                                  n_rounds))
        else:
            if not knn_ids.is_cuda or not knn_ids.is_contiguous() or knn_ids.element_size() != 4:
                raise RgError(RG_EINVAL, "knn_ids must be a contiguous 32-bit CUDA tensor")
            _check(lib().rg_index_build_roargraph(self._h, C.c_void_p(knn_ids.data_ptr()), 1, knn_ids.shape[0],
                                                  knn_ids.shape[1], M_sq, M_pjbp, L_pjpq, n_rounds))

    def graph_size(self):
        nnz, ep, md = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        _check(lib().rg_index_graph_size(self._h, C.byref(nnz), C.byref(ep), C.byref(md)))
        return dict(nnz=nnz.value, ep=ep.value, max_degree=md.value)

    def get_graph(self):
        """-> (ep, row_off u64[nd+1], nbrs u32[nnz]): what SaveProjectionGraph writes (src/index_bipartite.cpp:2606-2619)."""
        g = self.graph_size()
        row_off = np.empty(self.nd + 1, dtype=np.uint64)
        nbrs = np.empty(g["nnz"], dtype=np.uint32)
        _check(lib().rg_index_get_graph(self._h, _np_ptr(row_off), _np_ptr(nbrs)))
        return g["ep"], row_off, nbrs

    # -- SearchRoarGraph ---------------------------------------------------------------------------------------
    def search(self, queries: np.ndarray, k: int, L_pq: int):
        """Host buffers in, host buffers out (H2D + kernel + D2H inside the call).
        -> ids[nq,k] u32, dists[nq,k] f32 (internal distances), cmps[nq] u32, hops[nq] u32"""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise RgError(RG_EINVAL, "queries must be nq x dim (padded like the base)")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint32)
        dists = np.empty((nq, k), dtype=np.float32)
        cmps = np.empty(nq, dtype=np.uint32)
        hops = np.empty(nq, dtype=np.uint32)
        _check(lib().rg_search_roargraph(self._h, _np_ptr(q), nq, k, L_pq, _np_ptr(ids), _np_ptr(dists), _np_ptr(cmps),
                                         _np_ptr(hops)))
        return ids, dists, cmps, hops

    def search_into(self, q: np.ndarray, k: int, L_pq: int, ids, dists, cmps, hops) -> None:
        """Same call with caller-owned (e.g. pinned) host buffers; no allocation inside."""
        _check(lib().rg_search_roargraph(self._h, _np_ptr(q), q.shape[0], k, L_pq, _np_ptr(ids), _np_ptr(dists),
                                         _np_ptr(cmps) if cmps is not None else None,
                                         _np_ptr(hops) if hops is not None else None))

    def search_dev(self, d_q, k: int, L_pq: int, d_ids, d_dists, d_cmps=None, d_hops=None, stream=None) -> None:
        """Device tensors (torch, int32/float32 storage), asynchronous on `stream`."""
        _check(lib().rg_search_roargraph_dev(self._h, C.c_void_p(d_q.data_ptr()), d_q.shape[0], k, L_pq,
                                             C.c_void_p(d_ids.data_ptr()), C.c_void_p(d_dists.data_ptr()),
                                             C.c_void_p(d_cmps.data_ptr()) if d_cmps is not None else None,
                                             C.c_void_p(d_hops.data_ptr()) if d_hops is not None else None,
                                             _stream_ptr(stream)))

    def check(self, stream=None) -> None:
        _check(lib().rg_index_check(self._h, _stream_ptr(stream)))

    # -- SearchBipartiteGraph ----------------------------------------------------------------------------------
    def search_bipartite(self, queries: np.ndarray, k: int, L_pq: int, init_ids: np.ndarray | None = None,
                         n_init: int = 10, seed: int = 0):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint32)
        dists = np.empty((nq, k), dtype=np.float32)
        cmps = np.empty(nq, dtype=np.uint32)
        hops = np.empty(nq, dtype=np.uint32)
        ip = None
        if init_ids is not None:
            init_ids = np.ascontiguousarray(init_ids, dtype=np.uint32)
            n_init = init_ids.shape[1]
            ip = _np_ptr(init_ids)
        _check(lib().rg_search_bipartite(self._h, _np_ptr(q), nq, k, L_pq, ip, n_init, seed, _np_ptr(ids),
                                         _np_ptr(dists), _np_ptr(cmps), _np_ptr(hops)))
        return ids, dists, cmps, hops

    def search_bipartite_dev(self, d_q, k: int, L_pq: int, d_ids, d_dists, d_cmps=None, d_hops=None, d_init_ids=None,
                             n_init: int = 10, seed: int = 0, stream=None) -> None:
        """Device tensors, asynchronous on `stream`; d_init_ids: int32/uint32 [nq, n_init] or None (seeded draw)."""
        if d_init_ids is not None:
            n_init = d_init_ids.shape[1]
        _check(lib().rg_search_bipartite_dev(self._h, C.c_void_p(d_q.data_ptr()), d_q.shape[0], k, L_pq,
                                             C.c_void_p(d_init_ids.data_ptr()) if d_init_ids is not None else None,
                                             n_init, seed, C.c_void_p(d_ids.data_ptr()), C.c_void_p(d_dists.data_ptr()),
                                             C.c_void_p(d_cmps.data_ptr()) if d_cmps is not None else None,
                                             C.c_void_p(d_hops.data_ptr()) if d_hops is not None else None,
                                             _stream_ptr(stream)))

    # -- accounting --------------------------------------------------------------------------------------------
    def last_stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().rg_search_last_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(cmps=a.value, hops=b.value, adj_words=c.value)

    def last_launch(self):
        g, b, s, r = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        _check(lib().rg_search_last_launch(self._h, C.byref(g), C.byref(b), C.byref(s), C.byref(r)))
        return dict(grid=g.value, block=b.value, smem_bytes=s.value, rows_per_batch=r.value)


def knn(base: np.ndarray, queries: np.ndarray, K: int, metric: str | int, id_offset: int = 0):
    """exact_knn's contract on host buffers: -> ids[nq,K] int32, dists[nq,K] float32 ascending internal distances."""
    m = METRIC[metric] if isinstance(metric, str) else int(metric)
    b = np.ascontiguousarray(base, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    ids = np.empty((q.shape[0], K), dtype=np.int32)
    dists = np.empty((q.shape[0], K), dtype=np.float32)
    _check(lib().rg_knn(_np_ptr(b), b.shape[0], _np_ptr(q), q.shape[0], b.shape[1], K, m, id_offset, _np_ptr(ids),
                        _np_ptr(dists)))
    return ids, dists


def knn_dev(d_base, d_queries, K: int, metric: str | int, d_ids, d_dists, id_offset: int = 0, stream=None) -> None:
    m = METRIC[metric] if isinstance(metric, str) else int(metric)
    _check(lib().rg_knn_dev(C.c_void_p(d_base.data_ptr()), d_base.shape[0], C.c_void_p(d_queries.data_ptr()),
                            d_queries.shape[0], d_base.shape[1], K, m, id_offset, C.c_void_p(d_ids.data_ptr()),
                            C.c_void_p(d_dists.data_ptr()), _stream_ptr(stream)))


class KnnBase:
    """The base side of exact_knn prepared once (norms, normalised copy for cosine, bf16 operand copy) for any number of
    query batches — compute_groundtruth's base part kept in memory while the queries walk past it
    (compute_groundtruth.cpp:187-200, :396-416).  `base`: float32 numpy matrix (copied to HBM) or CUDA torch tensor (used in
    place; must outlive the object)."""

    def __init__(self, base, metric: str | int):
        self._h = C.c_void_p(0)
        m = METRIC[metric] if isinstance(metric, str) else int(metric)
        if isinstance(base, np.ndarray):
            b = np.ascontiguousarray(base, dtype=np.float32)
            self._keep = None
            self.nb, self.dim = b.shape
            _check(lib().rg_knn_base_create(C.byref(self._h), _np_ptr(b), self.nb, self.dim, m, 0))
        else:
            if not base.is_cuda or not base.is_contiguous() or base.dim() != 2 or str(base.dtype) != "torch.float32":
                raise RgError(RG_EINVAL, "base must be a contiguous float32 CUDA tensor")
            self._keep = base
            self.nb, self.dim = base.shape
            _check(lib().rg_knn_base_create(C.byref(self._h), C.c_void_p(base.data_ptr()), self.nb, self.dim, m, 1))

    def search_dev(self, d_queries, K: int, d_ids, d_dists, id_offset: int = 0, stream=None) -> None:
        _check(lib().rg_knn_base_search_dev(self._h, C.c_void_p(d_queries.data_ptr()), d_queries.shape[0], K, id_offset,
                                            C.c_void_p(d_ids.data_ptr()), C.c_void_p(d_dists.data_ptr()),
                                            _stream_ptr(stream)))

    def seed_dev(self, d_queries, K: int, rank: int, d_tau, stream=None) -> None:
        """d_tau[q] <- an upper bound of the distance of the rank-th best row of query q among this (shard of the) base's
        seed rows; the maximum over G shards with rank = ceil(K / G) bounds the global K-th best distance."""
        _check(lib().rg_knn_base_seed_dev(self._h, C.c_void_p(d_queries.data_ptr()), d_queries.shape[0], K, rank,
                                          C.c_void_p(d_tau.data_ptr()), _stream_ptr(stream)))

    def search_seeded_dev(self, d_queries, K: int, d_tau, d_ids, d_dists, id_offset: int = 0, stream=None,
                          bound_rows: int = 0) -> None:
        """search_dev with an externally agreed bound d_tau in force from row 0 (the per-shard lists may then hold fewer
        than K entries: id -1, distance +inf behind them)."""
        _check(lib().rg_knn_base_search_seeded_dev(self._h, C.c_void_p(d_queries.data_ptr()), d_queries.shape[0], K, id_offset,
                                                   C.c_void_p(d_tau.data_ptr()), bound_rows, C.c_void_p(d_ids.data_ptr()),
                                                   C.c_void_p(d_dists.data_ptr()), _stream_ptr(stream)))

    def search_resync_dev(self, d_queries, K: int, d_tau, d_ids, d_dists, rank: int, resync_rows: int, d_bound, reduce_max,
                          id_offset: int = 0, stream=None, bound_rows: int = 0) -> None:
        """search_seeded_dev with ONE re-sync of the bound inside the shard (rg_knn_base_search_resync_dev): at the first
        chunk boundary at or past `resync_rows` the shard's rank-th best exact distances so far are in d_bound[q_off:q_off+n]
        and reduce_max(q_off, n) must reduce that slice to the maximum over the shards on the current stream."""
        failure = []

        def hook(_user, q_off, n, _stream):
            try:
                reduce_max(int(q_off), int(n))
                return 0
            except Exception as e:  # an exception must not unwind through the C frames
                failure.append(e)
                return RG_ECUDA

        cb = KNN_BOUND_HOOK(hook)
        rc = lib().rg_knn_base_search_resync_dev(self._h, C.c_void_p(d_queries.data_ptr()), d_queries.shape[0], K, id_offset,
                                                 C.c_void_p(d_tau.data_ptr()), bound_rows, rank, resync_rows,
                                                 C.c_void_p(d_bound.data_ptr()), cb, None, C.c_void_p(d_ids.data_ptr()),
                                                 C.c_void_p(d_dists.data_ptr()), _stream_ptr(stream))
        if failure:
            raise failure[0]
        _check(rc)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().rg_knn_base_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def knn_profile(reset: bool = False) -> dict:
    """Device ms per phase of this thread's kNN calls (only filled under RG_KNN_TRACE=1)."""
    out = (C.c_double * 7)()
    _check(lib().rg_knn_profile(out))
    if reset:
        lib().rg_knn_profile_reset()
    keys = ("convert_base_ms", "convert_queries_ms", "seed_ms", "filter_ms", "merge_ms", "chunks", "overflow_retries")
    return {k: float(v) for k, v in zip(keys, out)}


def knn_trace(on: int) -> None:
    """1 / 0: per-phase timers of the kNN calls on / off (a diagnostic pass: they synchronise); -1: follow RG_KNN_TRACE."""
    _check(lib().rg_knn_trace(int(on)))


def knn_release_workspace() -> None:
    """Free the device scratch rg_knn / rg_knn_dev keep between calls."""
    _check(lib().rg_knn_release_workspace())


def topk_merge_dev(d_in_dists, d_in_ids, d_out_dists, d_out_ids, stream=None) -> None:
    """[G, nq, K] per-shard lists -> [nq, K] merged (compute_groundtruth.cpp:424-444)."""
    G, nq, K = d_in_dists.shape
    _check(lib().rg_topk_merge_dev(C.c_void_p(d_in_dists.data_ptr()), C.c_void_p(d_in_ids.data_ptr()), G, nq, K,
                                   C.c_void_p(d_out_dists.data_ptr()), C.c_void_p(d_out_ids.data_ptr()),
                                   _stream_ptr(stream)))


def topk_pack_dev(d_dists, d_ids, id_offset: int, d_packed, stream=None) -> None:
    """(dist, id)[...] -> one int64 word per entry (high = float bits, low = id + id_offset; 0xffffffff kept)."""
    _check(lib().rg_topk_pack_dev(C.c_void_p(d_dists.data_ptr()), C.c_void_p(d_ids.data_ptr()), d_dists.numel(), id_offset,
                                  C.c_void_p(d_packed.data_ptr()), _stream_ptr(stream)))


def topk_merge_packed_dev(d_packed, d_out_dists, d_out_ids, q_lo: int = 0, q_hi: int | None = None, stream=None) -> None:
    """[G, nq, K_in] packed words (an all-gather result) -> the K_out best per query into out[q_lo:q_hi]; out: [nq, K_out]."""
    G, nq, K_in = d_packed.shape
    K_out = d_out_dists.shape[1]
    _check(lib().rg_topk_merge_packed_dev(C.c_void_p(d_packed.data_ptr()), G, nq, K_in, K_out, q_lo,
                                          nq if q_hi is None else q_hi, C.c_void_p(d_out_dists.data_ptr()),
                                          C.c_void_p(d_out_ids.data_ptr()), _stream_ptr(stream)))


class Group:
    """Several GPUs of one box driven from THIS process (rg_group_*: one worker thread and one NCCL rank per GPU) — the
    single-process form the C++ CLIs use for BASELINE.json's multi-GPU configurations."""

    def __init__(self, n_gpus: int, dev_ids=None):
        self._h = C.c_void_p(0)
        arr = (C.c_int * n_gpus)(*dev_ids) if dev_ids is not None else None
        _check(lib().rg_group_create(C.byref(self._h), n_gpus, arr))
        self.n = n_gpus

    def knn(self, base: np.ndarray, queries: np.ndarray, K: int, metric: str | int):
        """exact_knn with the base sharded over the group's GPUs: -> ids[nq,K] int32, dists[nq,K] float32."""
        m = METRIC[metric] if isinstance(metric, str) else int(metric)
        b = np.ascontiguousarray(base, dtype=np.float32)
        q = np.ascontiguousarray(queries, dtype=np.float32)
        ids = np.empty((q.shape[0], K), dtype=np.int32)
        dists = np.empty((q.shape[0], K), dtype=np.float32)
        _check(lib().rg_group_knn(self._h, _np_ptr(b), b.shape[0], _np_ptr(q), q.shape[0], b.shape[1], K, m, _np_ptr(ids),
                                  _np_ptr(dists)))
        return ids, dists

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().rg_group_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupShardedIndex:
    """Base rows sharded over the GPUs of a Group, one RoarGraph per shard (rg_sharded_index_*)."""

    def __init__(self, group: Group, base: np.ndarray, metric: str | int):
        self._h = C.c_void_p(0)
        self.group = group
        m = METRIC[metric] if isinstance(metric, str) else int(metric)
        self.metric = m
        b = np.ascontiguousarray(base, dtype=np.float32)
        self.nd, self.dim = b.shape
        _check(lib().rg_sharded_index_create(group._h, C.byref(self._h), _np_ptr(b), self.nd, self.dim, m))

    def shard(self, s: int):
        """-> (RoarIndex view of shard s, first global row, rows)"""
        h, lo, n = C.c_void_p(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().rg_sharded_index_shard(self._h, s, C.byref(h), C.byref(lo), C.byref(n)))
        return RoarIndex._borrowed(h.value, n.value, self.dim, self.metric), lo.value, n.value

    def build_roargraph(self, train: np.ndarray, M_sq: int = 100, M_pjbp: int = 35, L_pjpq: int = 500) -> None:
        t = np.ascontiguousarray(train, dtype=np.float32)
        _check(lib().rg_sharded_index_build_roargraph(self._h, _np_ptr(t), t.shape[0], M_sq, M_pjbp, L_pjpq))

    def search(self, queries: np.ndarray, k: int, L_pq: int):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        dists = np.empty((q.shape[0], k), dtype=np.float32)
        _check(lib().rg_sharded_search_roargraph(self._h, _np_ptr(q), q.shape[0], k, L_pq, _np_ptr(ids), _np_ptr(dists)))
        return ids, dists

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().rg_sharded_index_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
