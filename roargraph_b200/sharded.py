"""Base-sharded search across the GPUs of one box: one process per GPU, every rank owns a slice of the base rows and
its own RoarGraph over that slice, every query runs on every shard, and ONE all-gather of the per-shard top-k
(distance, global id) lists followed by a k-way merge kernel yields the final top-k on every rank (SURVEY.md 8(e),
BASELINE.json north_star).  The merge is the std::sort merge of compute_groundtruth.cpp:424-444 applied to search
results: ascending (distance, id).

torch.distributed is the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests); the merge itself is the
CUDA kernel rg_topk_merge_dev — there is no CPU merge in the product.
"""
from __future__ import annotations

import numpy as np

import roargraph_b200 as rg


def shard_range(nb: int, shard: int, n_shards: int):
    """Rows [lo, hi) of shard `shard`: the same split compute_groundtruth uses for base parts, by equal counts."""
    return shard * nb // n_shards, (shard + 1) * nb // n_shards


def gather_topk(local_dists, local_ids, group=None, out=None):
    """all-gather [nq, k] per-shard lists into [G, nq, k] tensors (same device/backend as the inputs).
    `out` = optional (dists [G*nq, k], ids [G*nq, k]) buffers to gather into."""
    import torch
    import torch.distributed as dist

    G = dist.get_world_size(group)
    nq, k = local_dists.shape
    # concatenated along dim 0 (the layout both NCCL and gloo accept), viewed as [G][nq][k]
    if out is not None:
        out_d, out_i = out
    else:
        out_d = torch.empty((G * nq, k), dtype=local_dists.dtype, device=local_dists.device)
        out_i = torch.empty((G * nq, k), dtype=local_ids.dtype, device=local_ids.device)
    dist.all_gather_into_tensor(out_d, local_dists.contiguous(), group=group)
    dist.all_gather_into_tensor(out_i, local_ids.contiguous(), group=group)
    return out_d.view(G, nq, k), out_i.view(G, nq, k)


class ShardedRoarIndex:
    """A RoarIndex over base rows [row_lo, row_lo + nd) of a larger base; search() returns GLOBAL ids merged over all
    ranks of `group` (identical on every rank).  With a single rank it degenerates to the local search."""

    def __init__(self, index: "rg.RoarIndex", row_lo: int = 0, group=None):
        import torch.distributed as dist

        self.index = index
        self.row_lo = int(row_lo)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._buf = {}

    def _scratch(self, nq, k, device):
        import torch

        key = (nq, k)
        if key not in self._buf:
            G = self.world
            self._buf[key] = (torch.empty((nq, k), dtype=torch.int32, device=device),
                              torch.empty((nq, k), dtype=torch.float32, device=device),
                              torch.empty((G * nq, k), dtype=torch.float32, device=device),
                              torch.empty((G * nq, k), dtype=torch.int32, device=device))
        return self._buf[key]

    def search_dev(self, d_q, k: int, L_pq: int, out_ids, out_dists, d_cmps=None, d_hops=None) -> int:
        """Device tensors in/out on torch's current stream.  Returns the number of kernels launched."""
        import torch

        stream = torch.cuda.current_stream()
        if self.world == 1:
            self.index.search_dev(d_q, k, L_pq, out_ids, out_dists, d_cmps, d_hops, stream=stream)
            if self.row_lo:
                out_ids.add_(self.row_lo)
                return 2
            return 1
        loc_i, loc_d, g_d, g_i = self._scratch(d_q.shape[0], k, d_q.device)
        self.index.search_dev(d_q, k, L_pq, loc_i, loc_d, d_cmps, d_hops, stream=stream)
        loc_i.add_(self.row_lo)  # local -> global ids, like start_id in compute_groundtruth.cpp:413
        all_d, all_i = gather_topk(loc_d, loc_i, self.group, out=(g_d, g_i))
        rg.topk_merge_dev(all_d, all_i, out_dists, out_ids, stream=stream)
        return 3

    def _e2e_buffers(self, nq, dim, k, device):
        import torch

        key = ("e2e", nq, dim, k)
        if key not in self._buf:
            self._buf[key] = dict(
                d_q=torch.empty((nq, dim), dtype=torch.float32, device=device),
                d_i=torch.empty((nq, k), dtype=torch.int32, device=device),
                d_d=torch.empty((nq, k), dtype=torch.float32, device=device),
                h_i=torch.empty((nq, k), dtype=torch.int32).pin_memory(),
                h_d=torch.empty((nq, k), dtype=torch.float32).pin_memory())
        return self._buf[key]

    def search(self, queries: np.ndarray, k: int, L_pq: int):
        """Host buffers in (ideally pinned), host numpy out: H2D, per-shard search, all-gather, merge, D2H.  All device
        and pinned staging buffers are cached per shape, every copy is asynchronous on torch's current stream and the
        call synchronises once."""
        import torch

        dev = torch.device("cuda", torch.cuda.current_device())
        q = torch.from_numpy(queries)
        nq, dim = q.shape
        b = self._e2e_buffers(nq, dim, k, dev)
        b["d_q"].copy_(q, non_blocking=True)
        self.search_dev(b["d_q"], k, L_pq, b["d_i"], b["d_d"])
        b["h_i"].copy_(b["d_i"], non_blocking=True)
        b["h_d"].copy_(b["d_d"], non_blocking=True)
        self.index.check(torch.cuda.current_stream())  # synchronises the stream, raises the reference's errors
        return b["h_i"].numpy().view(np.uint32), b["h_d"].numpy()


class ReplicatedRoarIndex:
    """Query-sharded mode ("replicas"): every rank holds the FULL base and the same RoarGraph, owns a
    slice of the query batch, and the per-rank result blocks are collected with one all-gather — no merge step, and the
    ids are those of the single-GPU search.  The comparison point for the base-sharded contract: G x the memory, the
    same work per query, throughput ~ linear in G; base-sharding does G x the work per query at a smaller L_pq.

    Every rank must own the same number of queries (all_gather_into_tensor), so the batch is padded by repeating the
    last query of a short slice; search_dev returns only the real rows."""

    def __init__(self, index, group=None, local_search=None):
        import torch.distributed as dist

        self.index = index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._local = local_search  # tests inject a checker; default = the CUDA search on torch's current stream
        self._buf = {}

    def _search_local(self, d_q, k, L_pq, out_ids, out_dists):
        if self._local is not None:
            return self._local(d_q, k, L_pq, out_ids, out_dists)
        import torch
        self.index.search_dev(d_q, k, L_pq, out_ids, out_dists, None, None, stream=torch.cuda.current_stream())

    def search_dev(self, d_q_all, k: int, L_pq: int, out_ids, out_dists) -> int:
        """d_q_all: the WHOLE batch [nq, dim] on this rank's device (only this rank's slice is read);
        out_ids / out_dists: [nq, k], complete on every rank afterwards.  Returns the number of kernels launched."""
        import torch
        import torch.distributed as dist

        nq = d_q_all.shape[0]
        if self.world == 1:
            self._search_local(d_q_all, k, L_pq, out_ids, out_dists)
            return 1
        per = -(-nq // self.world)
        lo = min(self.rank * per, nq)
        hi = min(lo + per, nq)
        key = (per, k, d_q_all.device)
        if key not in self._buf:
            self._buf[key] = (torch.empty((per, d_q_all.shape[1]), dtype=d_q_all.dtype, device=d_q_all.device),
                              torch.zeros((per, k), dtype=out_ids.dtype, device=d_q_all.device),
                              torch.zeros((per, k), dtype=out_dists.dtype, device=d_q_all.device),
                              torch.empty((self.world * per, k), dtype=out_ids.dtype, device=d_q_all.device),
                              torch.empty((self.world * per, k), dtype=out_dists.dtype, device=d_q_all.device))
        q_loc, loc_i, loc_d, all_i, all_d = self._buf[key]
        launches = 0
        if hi - lo == per:
            self._search_local(d_q_all[lo:hi], k, L_pq, loc_i, loc_d)  # rows of a row-major batch: contiguous, in place
            launches = 1
        elif hi > lo:
            q_loc[:hi - lo].copy_(d_q_all[lo:hi])
            q_loc[hi - lo:].copy_(d_q_all[hi - 1:hi].expand(per - (hi - lo), -1))
            self._search_local(q_loc, k, L_pq, loc_i, loc_d)
            launches = 1
        dist.all_gather_into_tensor(all_i, loc_i, group=self.group)
        dist.all_gather_into_tensor(all_d, loc_d, group=self.group)
        out_ids.copy_(all_i[:nq])   # rank r's block starts at r * per: contiguous in query order, padding only at the end
        out_dists.copy_(all_d[:nq])
        return launches


def sharded_knn_dev(d_base_shard, d_queries, K: int, metric, row_lo: int, group=None, batch: int = 1 << 20,
                    local_knn=None, merge=None):
    """Brute-force kNN with the BASE sharded over the ranks of `group` (BASELINE.json configs[3]): every rank scores all
    queries against its base rows [row_lo, row_lo + nb_local) — the parts of compute_groundtruth.cpp:396-416, ids offset
    by the part's start (:413) — then, per query batch, ONE all-gather of the per-shard (distance, id)[batch][K] lists
    and the k-way merge kernel (the std::sort merge of :424-444) give the global top-K on every rank.

    d_base_shard / d_queries: device tensors (fp32, row-major).  Returns (ids int32 [nq, K], dists fp32 [nq, K]) on the
    device: internal distances ascending (-IP for "ip", squared L2 for "l2"/"cosine"), global ids.
    `local_knn` / `merge` default to the CUDA entry points (rg.knn_dev, rg.topk_merge_dev); the CPU tests inject
    checkers there to exercise the batching, offsets and collective layout over gloo."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    nq = d_queries.shape[0]
    dev = d_queries.device
    if local_knn is None:
        def local_knn(b, q, k, m, ids, dists, id_offset):
            rg.knn_dev(b, q, k, m, ids, dists, id_offset=id_offset, stream=torch.cuda.current_stream())
    if merge is None:
        def merge(all_d, all_i, out_d, out_i):
            rg.topk_merge_dev(all_d, all_i, out_d, out_i, stream=torch.cuda.current_stream())
    loc_i = torch.empty((nq, K), dtype=torch.int32, device=dev)
    loc_d = torch.empty((nq, K), dtype=torch.float32, device=dev)
    local_knn(d_base_shard, d_queries, K, metric, loc_i, loc_d, int(row_lo))
    if world == 1:
        return loc_i, loc_d
    out_i = torch.empty_like(loc_i)
    out_d = torch.empty_like(loc_d)
    for lo in range(0, nq, batch):
        hi = min(nq, lo + batch)
        all_d, all_i = gather_topk(loc_d[lo:hi], loc_i[lo:hi], group)
        merge(all_d, all_i, out_d[lo:hi], out_i[lo:hi])
    return out_i, out_d


def query_sharded_knn_dev(d_base_full, d_queries, K: int, metric, group=None, local_knn=None):
    """Brute-force kNN with the QUERIES sharded and the base replicated — the comparison point SURVEY.md 8(e) names for
    the base-sharded contract above: rank r scores queries [r * per, (r + 1) * per) against ALL base rows and one
    all-gather collects the per-rank result blocks; no merge step, and every rank's tensor-core filter streams the whole
    base (the base-sharded form runs G times as many, G times smaller filter passes).  Needs the full base on every GPU.

    Returns (ids int32 [nq, K], dists fp32 [nq, K]) on the device, complete on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    nq = d_queries.shape[0]
    dev = d_queries.device
    if local_knn is None:
        def local_knn(b, q, k, m, ids, dists, id_offset):
            rg.knn_dev(b, q, k, m, ids, dists, id_offset=id_offset, stream=torch.cuda.current_stream())
    per = -(-nq // world)
    lo = min(rank * per, nq)
    hi = min(lo + per, nq)
    loc_i = torch.zeros((per, K), dtype=torch.int32, device=dev)
    loc_d = torch.zeros((per, K), dtype=torch.float32, device=dev)
    if hi > lo:
        local_knn(d_base_full, d_queries[lo:hi], K, metric, loc_i[:hi - lo], loc_d[:hi - lo], 0)
    if world == 1:
        return loc_i[:nq], loc_d[:nq]
    all_i = torch.empty((world * per, K), dtype=torch.int32, device=dev)
    all_d = torch.empty((world * per, K), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(all_i, loc_i, group=group)
    dist.all_gather_into_tensor(all_d, loc_d, group=group)
    return all_i[:nq], all_d[:nq]  # rank r's block starts at r * per: query order, padding only behind the last query
