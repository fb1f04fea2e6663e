"""ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  Nothing under roargraph_b200/ imports it (tests/test_no_oracle_in_product.py enforces that).

* ``Oracle``    -> oracle/liboracle.so        (plain-C restatement, oracle/roar_oracle.c)
* ``Reference`` -> oracle/_ref/libref_roargraph.so (the reference's own code compiled unmodified, oracle/ref_capi.cpp)
* ``Reference(detrand=True)`` -> oracle/_ref/libref_roargraph_detrand.so: the same translation units with
  oracle/shims/fixed_random_device.h pre-included, so that SearchBipartiteGraph's entry points are reproducible
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
METRIC = {"l2": 0, "ip": 1, "cosine": 4}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (oracle/Makefile).  `ref` also builds oracle/_ref when /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []))


def have_reference() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_roargraph.so"))


def have_detrand_reference() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_roargraph_detrand.so"))


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_distance.restype = C.c_float
        L.orc_distance.argtypes = [C.c_int, _f32p, _f32p, C.c_uint]
        L.orc_normalize_rows.argtypes = [_f32p, C.c_uint64, C.c_uint]
        L.orc_search_roargraph.restype = C.c_int
        L.orc_search_roargraph.argtypes = [_f32p, C.c_uint64, C.c_uint, _u64p, _u32p, C.c_uint32, C.c_int, _f32p,
                                           C.c_uint64, C.c_uint, C.c_uint, _u32p, _f32p, _u32p, _u32p, _u64p, _i32p,
                                           C.c_int]
        L.orc_search_bipartite.restype = C.c_int
        L.orc_search_bipartite.argtypes = [_f32p, C.c_uint64, C.c_uint, _u64p, _u32p, _u32p, C.c_uint, C.c_int, _f32p,
                                           C.c_uint64, C.c_uint, C.c_uint, _u32p, _f32p, _u32p, _u32p, _i32p, C.c_int]
        L.orc_exact_knn.restype = C.c_int
        L.orc_exact_knn.argtypes = [C.c_uint, C.c_uint, _i32p, _f32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_int,
                                    C.c_int]
        L.orc_num_procs.restype = C.c_int

    def num_procs(self) -> int:
        return int(self.lib.orc_num_procs())

    def distance(self, metric: str, a, b) -> float:
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        return float(self.lib.orc_distance(METRIC[metric], a, b, a.shape[0]))

    def normalize_rows(self, x: np.ndarray) -> np.ndarray:
        x = np.array(x, dtype=np.float32, order="C", copy=True)
        self.lib.orc_normalize_rows(x, x.shape[0], x.shape[1])
        return x

    def search_roargraph(self, base, row_off, nbrs, ep, metric, queries, k, L, threads=0):
        """base/queries: padded (x8) row-major float32, already normalised for cosine.
        -> dict(ids, dists, cmps, hops, adj_words, status)"""
        nq = queries.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint32)
        dists = np.zeros((nq, k), dtype=np.float32)
        cmps = np.zeros(nq, dtype=np.uint32)
        hops = np.zeros(nq, dtype=np.uint32)
        adj = np.zeros(nq, dtype=np.uint64)
        status = np.zeros(nq, dtype=np.int32)
        nfail = self.lib.orc_search_roargraph(base, base.shape[0], base.shape[1], row_off, nbrs, int(ep),
                                              METRIC[metric], queries, nq, k, L, ids, dists, cmps, hops, adj, status,
                                              threads)
        return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, adj_words=adj, status=status, nfail=nfail)

    def search_bipartite(self, base, row_off, nbrs, init_ids, metric, queries, k, L, threads=0):
        nq = queries.shape[0]
        init_ids = np.ascontiguousarray(init_ids, dtype=np.uint32)
        ids = np.zeros((nq, k), dtype=np.uint32)
        dists = np.zeros((nq, k), dtype=np.float32)
        cmps = np.zeros(nq, dtype=np.uint32)
        hops = np.zeros(nq, dtype=np.uint32)
        status = np.zeros(nq, dtype=np.int32)
        nfail = self.lib.orc_search_bipartite(base, base.shape[0], base.shape[1], row_off, nbrs, init_ids,
                                              init_ids.shape[1], METRIC[metric], queries, nq, k, L, ids, dists, cmps,
                                              hops, status, threads)
        return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, status=status, nfail=nfail)

    def exact_knn(self, base, queries, K, metric, threads=0):
        """Unpadded row-major inputs, internal distances (IP negative), as exact_knn's contract."""
        base = np.ascontiguousarray(base, dtype=np.float32)
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = queries.shape[0]
        ids = np.zeros((nq, K), dtype=np.int32)
        dists = np.zeros((nq, K), dtype=np.float32)
        rc = self.lib.orc_exact_knn(base.shape[1], K, ids, dists, base.shape[0], base, nq, queries, METRIC[metric],
                                    threads)
        if rc != 0:
            raise RuntimeError("orc_exact_knn failed")
        return ids, dists


class Reference:
    """The unmodified reference (prebuilt oracle/_ref/libref_roargraph.so; built from /root/reference when present)."""

    def __init__(self, detrand: bool = False):
        path = os.path.join(HERE, "_ref", "libref_roargraph_detrand.so" if detrand else "libref_roargraph.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle ref` where /root/reference is mounted")
        self.lib = L = C.CDLL(path)
        self.detrand = detrand
        L.ref_has_fixed_random_device.restype = C.c_int
        assert bool(L.ref_has_fixed_random_device()) == detrand
        L.ref_bipartite_load.restype = C.c_void_p
        L.ref_bipartite_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint, C.c_uint64, C.c_uint, C.c_int]
        L.ref_bipartite_entry_points.argtypes = [C.c_uint32, C.c_uint32, _u32p]
        L.ref_search_bipartite.restype = C.c_int
        L.ref_search_bipartite.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.c_uint32, _u32p,
                                           _f32p, _u32p, _u32p, _u32p, C.POINTER(C.c_double)]
        L.ref_last_error.restype = C.c_char_p
        L.ref_num_procs.restype = C.c_int
        L.ref_distance.restype = C.c_float
        L.ref_distance.argtypes = [C.c_int, _f32p, _f32p, C.c_uint]
        L.ref_build_roargraph.restype = C.c_int
        L.ref_build_roargraph.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_uint, C.c_uint, C.c_uint,
                                          C.c_uint, C.c_int]
        L.ref_index_load.restype = C.c_void_p
        L.ref_index_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint, C.c_uint, C.c_int]
        L.ref_index_from_memory.restype = C.c_void_p
        L.ref_index_from_memory.argtypes = [_f32p, C.c_uint64, C.c_uint, C.c_int, C.c_char_p, C.c_uint, C.c_int]
        L.ref_index_need_normalize.argtypes = [C.c_void_p]
        L.ref_normalize_rows.argtypes = [_f32p, C.c_uint64, C.c_uint]
        L.ref_search_roargraph.restype = C.c_int
        L.ref_search_roargraph.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.c_int, _u32p,
                                           _f32p, _u32p, _u32p, C.POINTER(C.c_double)]
        L.ref_index_free.argtypes = [C.c_void_p]
        L.ref_index_base.restype = C.c_void_p
        L.ref_index_base.argtypes = [C.c_void_p]

    def err(self) -> str:
        return self.lib.ref_last_error().decode()

    def num_procs(self) -> int:
        return int(self.lib.ref_num_procs())

    def distance(self, metric: str, a, b) -> float:
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        return float(self.lib.ref_distance(METRIC[metric], a, b, a.shape[0]))

    def normalize_rows(self, x: np.ndarray) -> np.ndarray:
        x = np.array(x, dtype=np.float32, order="C", copy=True)
        self.lib.ref_normalize_rows(x, x.shape[0], x.shape[1])
        return x

    def build_roargraph(self, base_fbin, knn_file, out_index, metric, M_sq, M_pjbp, L_pjpq, threads, quiet=True):
        rc = self.lib.ref_build_roargraph(base_fbin.encode(), knn_file.encode(), out_index.encode(), METRIC[metric],
                                          M_sq, M_pjbp, L_pjpq, threads, int(quiet))
        if rc != 0:
            raise RuntimeError("reference build failed: " + self.err())

    def index_load(self, base_fbin, index_file, metric, padded_dim, threads, quiet=True):
        h = self.lib.ref_index_load(base_fbin.encode(), index_file.encode(), METRIC[metric], padded_dim, threads,
                                    int(quiet))
        if not h:
            raise RuntimeError("reference index load failed: " + self.err())
        return h

    def index_from_memory(self, base, index_file, metric, threads, quiet=True):
        """The reference index over a base array already in memory (padded x8, normalised for cosine); `base` must
        outlive the handle."""
        h = self.lib.ref_index_from_memory(base, base.shape[0], base.shape[1], METRIC[metric], index_file.encode(),
                                           threads, int(quiet))
        if not h:
            raise RuntimeError("reference index load failed: " + self.err())
        return h

    def index_base(self, h, nd, dim) -> np.ndarray:
        """View of the reference's own (padded, maybe normalised) base array."""
        ptr = self.lib.ref_index_base(h)
        buf = (C.c_float * (nd * dim)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(nd, dim)

    def search_roargraph(self, h, queries, k, L, threads, warmup=False):
        nq = queries.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint32)
        dists = np.zeros((nq, k), dtype=np.float32)
        cmps = np.zeros(nq, dtype=np.uint32)
        hops = np.zeros(nq, dtype=np.uint32)
        wall = C.c_double(0.0)
        rc = self.lib.ref_search_roargraph(h, queries, nq, k, L, threads, int(warmup), ids, dists, cmps, hops,
                                           C.byref(wall))
        if rc != 0:
            raise RuntimeError(self.err())
        return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, wall_s=wall.value)

    def index_free(self, h):
        self.lib.ref_index_free(h)

    # -- SearchBipartiteGraph (src/index_bipartite.cpp:282-356) ----------------------------------------------------
    def bipartite_load(self, base_fbin, bipartite_index_file, metric, padded_dim, n_sq, threads, quiet=True):
        """tests/test_search_bipartite.cpp:158-168: ctor -> LoadSearchNeededData -> Load -> InitVisitedListPool."""
        h = self.lib.ref_bipartite_load(base_fbin.encode(), bipartite_index_file.encode(), METRIC[metric], padded_dim,
                                        n_sq, threads, int(quiet))
        if not h:
            raise RuntimeError("reference bipartite index load failed: " + self.err())
        return h

    def bipartite_entry_points(self, seed_word: int, nd: int) -> np.ndarray:
        """The 10 ids mt19937(seed_word) + uniform_int_distribution(0, nd - 1) yield (:287-294), libstdc++'s own."""
        out = np.zeros(10, dtype=np.uint32)
        self.lib.ref_bipartite_entry_points(seed_word & 0xffffffff, nd, out)
        return out

    def search_bipartite(self, h, queries, k, L, threads, seed_base=0):
        """The reference's SearchBipartiteGraph over a batch.  With the detrand library query i draws its entry points
        from the seed word seed_base + i and `init_ids` returns them; with the plain library they are unseeded."""
        nq = queries.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint32)
        dists = np.zeros((nq, k), dtype=np.float32)
        cmps = np.zeros(nq, dtype=np.uint32)
        hops = np.zeros(nq, dtype=np.uint32)
        init = np.zeros((nq, 10), dtype=np.uint32)
        wall = C.c_double(0.0)
        rc = self.lib.ref_search_bipartite(h, queries, nq, k, L, threads, seed_base & 0xffffffff, ids, dists, cmps, hops,
                                           init, C.byref(wall))
        if rc != 0:
            raise RuntimeError(self.err())
        return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, init_ids=init if self.detrand else None,
                    wall_s=wall.value)


def run_ref_build_roargraph(base_fbin, train_fbin, knn_file, out_index, dist, M_sq, M_pjbp, L_pjpq, threads, quiet=True,
                            retries=3):
    """Run the reference's own build CLI (tests/test_build_roargraph.cpp compiled unmodified) in a SUBPROCESS.
    Its multi-threaded build reads neighbour vectors that other threads are growing ("benign-by-design"
    races, e.g. src/index_bipartite.cpp:1325 vs :1365); under 16 threads that occasionally corrupts the heap, so the
    build is isolated from the test process and retried if it dies."""
    exe = os.path.join(HERE, "_ref", "ref_test_build_roargraph")
    cmd = [exe, "--data_type", "float", "--dist", dist, "--base_data_path", base_fbin, "--sampled_query_data_path",
           train_fbin, "--projection_index_save_path", out_index, "--learn_base_nn_path", knn_file, "--M_sq", str(M_sq),
           "--M_pjbp", str(M_pjbp), "--L_pjpq", str(L_pjpq), "-T", str(threads)]
    last = None
    for _ in range(retries):
        if os.path.exists(out_index):
            os.remove(out_index)
        p = subprocess.run(cmd, stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.STDOUT if quiet else None)
        if p.returncode == 0 and os.path.exists(out_index):
            return
        last = p.returncode
    raise RuntimeError(f"reference build CLI failed (rc={last})")


def run_ref_compute_groundtruth(base_fbin, query_fbin, gt_file, K, dist_fn, quiet=True):
    """Run the reference's compute_groundtruth binary (compiled unmodified; MKL via torch's libtorch_cpu.so)."""
    exe = os.path.join(HERE, "_ref", "ref_compute_groundtruth")
    cmd = [exe, "--data_type", "float", "--dist_fn", dist_fn, "--base_file", base_fbin, "--query_file", query_fbin,
           "--gt_file", gt_file, "--K", str(K)]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL if quiet else None)
