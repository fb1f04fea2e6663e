"""Multi-GPU forms of the two hot paths across the GPUs of one box, one process per GPU (torch.distributed is the
plumbing: NCCL over NVLink on the GPU box, gloo in the CPU tests).

Base-sharded (the contract of BASELINE.json's north_star): every rank owns a slice of the base rows — and,
for the graph search, its own RoarGraph over that slice; every query runs on every shard; per exchange step ONE all-gather
of the per-shard top-k blocks, each (distance, global id) pair packed into one 8-byte word (rg_topk_pack_dev), followed by
the k-way merge kernel (rg_topk_merge_packed_dev) = the std::sort merge of compute_groundtruth.cpp:424-444: ascending
(distance, id), ids offset by the shard's first row like start_id (:413).  There is no CPU merge in the product.

Query-sharded replicas (the comparison point 8(e) names): the full base / index on every rank, the queries split, one
all-gather of the result blocks, no merge.
"""
from __future__ import annotations

import numpy as np

import roargraph_b200 as rg

_PACKED_MISSING = (0x7f800000 << 32) | 0xffffffff  # (+inf, no id): sorts after every real entry
NO_ID = -1  # int32 view of the 0xffffffff "no result" sentinel the search kernel writes for a short result list


def shard_range(nb: int, shard: int, n_shards: int):
    """Rows [lo, hi) of shard `shard`: the same split compute_groundtruth uses for base parts, by equal counts."""
    return shard * nb // n_shards, (shard + 1) * nb // n_shards


# kernels of this package behind one local graph search (rg_search_roargraph_dev / rg_search_bipartite_dev): tags_clear_kernel,
# tags_reset_kernel (visited-tag maintenance in front of every launch) and beam_search_kernel — what an ncu launch list of
# a step shows (profiles/r2_launches_10M.csv); pack and merge are one kernel each
SEARCH_KERNELS = 3


def _world(group=None):
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_topk(local_dists, local_ids, group=None, out=None):
    """all-gather [nq, k] per-shard lists into [G, nq, k] tensors (same device/backend as the inputs): the unpacked
    two-collective form, kept for CPU tensors (gloo tests); the CUDA paths below exchange packed words in one collective.
    `out` = optional (dists [G*nq, k], ids [G*nq, k]) buffers to gather into."""
    import torch
    import torch.distributed as dist

    G = dist.get_world_size(group)
    nq, k = local_dists.shape
    if out is not None:
        out_d, out_i = out
    else:
        out_d = torch.empty((G * nq, k), dtype=local_dists.dtype, device=local_dists.device)
        out_i = torch.empty((G * nq, k), dtype=local_ids.dtype, device=local_ids.device)
    dist.all_gather_into_tensor(out_d, local_dists.contiguous(), group=group)
    dist.all_gather_into_tensor(out_i, local_ids.contiguous(), group=group)
    return out_d.view(G, nq, k), out_i.view(G, nq, k)


def pack_topk(dists, ids, id_offset: int, out=None, stream=None):
    """(distance, id)[...] -> one int64 word per entry: high word = the float's bits, low word = id + id_offset, the
    sentinel id kept (rg_topk_pack_dev on CUDA tensors; the same arithmetic in torch for CPU tensors in the gloo tests)."""
    import torch

    if out is None:
        out = torch.empty(dists.shape, dtype=torch.int64, device=dists.device)
    if dists.is_cuda:
        rg.topk_pack_dev(dists, ids, id_offset, out, stream=stream if stream is not None else torch.cuda.current_stream())
        return out
    i64 = ids.to(torch.int64) & 0xffffffff
    glob = torch.where(i64 == 0xffffffff, i64, (i64 + id_offset) & 0xffffffff)
    out.copy_((dists.contiguous().view(torch.int32).to(torch.int64) << 32) | glob)
    return out


def merge_packed(packed, out_dists, out_ids, q_lo: int = 0, q_hi: int | None = None, stream=None):
    """[G, nq, K_in] packed words -> the K_out best (distance asc, id asc) per query into out[q_lo:q_hi]
    (rg_topk_merge_packed_dev; a numpy checker of the same order for CPU tensors in the gloo tests)."""
    import torch

    G, nq, K_in = packed.shape
    q_hi = nq if q_hi is None else q_hi
    if packed.is_cuda:
        rg.topk_merge_packed_dev(packed, out_dists, out_ids, q_lo, q_hi,
                                 stream=stream if stream is not None else torch.cuda.current_stream())
        return
    K_out = out_dists.shape[1]
    w = packed.numpy()[:, q_lo:q_hi].transpose(1, 0, 2).reshape(q_hi - q_lo, G * K_in)
    d = (w >> 32).astype(np.int32).view(np.float32)
    i = (w & 0xffffffff).astype(np.uint32)
    d = np.where(i == 0xffffffff, np.float32(np.inf), d)
    order = np.lexsort((i, d), axis=1)[:, :K_out]
    out_dists[q_lo:q_hi] = torch.from_numpy(np.take_along_axis(d, order, 1))
    out_ids[q_lo:q_hi] = torch.from_numpy(np.take_along_axis(i, order, 1).view(np.int32))


def raise_on_all_ranks(local_error: Exception | None, device, group=None):
    """A rank whose local search failed ("not enough results") must not leave its peers waiting in the next collective:
    the error flag is reduced over the ranks and every rank raises."""
    import torch
    import torch.distributed as dist

    world, _ = _world(group)
    flag = torch.tensor([1 if local_error is not None else 0], dtype=torch.int32, device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if int(flag.item()):
        if local_error is not None:
            raise local_error
        raise rg.RgError(rg.RG_ENOTENOUGH, "not enough results on another shard")


class ShardedRoarIndex:
    """A RoarIndex over base rows [row_lo, row_lo + nd) of a larger base; search() returns GLOBAL ids merged over all
    ranks of `group` (identical on every rank).  With a single rank it degenerates to the local search.

    `local_search(d_q, k, L_pq, out_ids, out_dists)` replaces the CUDA search in the CPU tests."""

    def __init__(self, index: "rg.RoarIndex", row_lo: int = 0, group=None, local_search=None):
        self.index = index
        self.row_lo = int(row_lo)
        self.group = group
        self.world, self.rank = _world(group)
        self._local = local_search
        self._buf = {}

    def _scratch(self, nq, k, device):
        import torch

        key = (nq, k, str(device))
        if key not in self._buf:
            G = self.world
            self._buf[key] = (torch.empty((nq, k), dtype=torch.int32, device=device),
                              torch.empty((nq, k), dtype=torch.float32, device=device),
                              torch.empty((nq, k), dtype=torch.int64, device=device),
                              torch.empty((G * nq, k), dtype=torch.int64, device=device))
        return self._buf[key]

    def _search_local(self, d_q, k, L_pq, out_ids, out_dists, d_cmps, d_hops):
        if self._local is not None:
            return self._local(d_q, k, L_pq, out_ids, out_dists)
        import torch
        self.index.search_dev(d_q, k, L_pq, out_ids, out_dists, d_cmps, d_hops, stream=torch.cuda.current_stream())

    def search_dev(self, d_q, k: int, L_pq: int, out_ids, out_dists, d_cmps=None, d_hops=None, k_shard: int | None = None) -> int:
        """Device tensors in/out on torch's current stream.  Every shard returns its k_shard best (default k; a shard
        needs at most its share of the global top-k, so k_shard < k with L_pq >= k_shard is allowed) and the merge keeps
        the k best of the G * k_shard gathered entries.  Returns the number of this package's kernels launched."""
        import torch.distributed as dist

        ks = k if k_shard is None else int(k_shard)
        if self.world == 1:
            if ks == k:
                self._search_local(d_q, k, L_pq, out_ids, out_dists, d_cmps, d_hops)
                if self.row_lo:
                    out_ids.add_((out_ids != NO_ID).to(out_ids.dtype) * self.row_lo)  # the sentinel stays what it is
                return SEARCH_KERNELS
        loc_i, loc_d, packed, gathered = self._scratch(d_q.shape[0], ks, d_q.device)
        self._search_local(d_q, ks, L_pq, loc_i, loc_d, d_cmps, d_hops)
        pack_topk(loc_d, loc_i, self.row_lo, out=packed)  # local -> global ids, like start_id in compute_groundtruth.cpp:413
        if self.world > 1:
            dist.all_gather_into_tensor(gathered, packed, group=self.group)  # the ONE exchange step
        else:
            gathered.copy_(packed)
        merge_packed(gathered.view(self.world, d_q.shape[0], ks), out_dists, out_ids)
        return SEARCH_KERNELS + 2

    def _e2e_buffers(self, nq, dim, k, device):
        import torch

        key = ("e2e", nq, dim, k)
        if key not in self._buf:
            per = -(-nq // self.world)
            self._buf[key] = dict(
                d_q=torch.empty((self.world * per, dim), dtype=torch.float32, device=device),
                d_mine=torch.zeros((per, dim), dtype=torch.float32, device=device),
                d_i=torch.empty((nq, k), dtype=torch.int32, device=device),
                d_d=torch.empty((nq, k), dtype=torch.float32, device=device),
                h_i=torch.empty((nq, k), dtype=torch.int32).pin_memory(),
                h_d=torch.empty((nq, k), dtype=torch.float32).pin_memory())
        return self._buf[key]

    def search(self, queries: np.ndarray, k: int, L_pq: int, k_shard: int | None = None):
        """Host buffers in (ideally pinned), host numpy out.  Every rank holds the same host batch; rank r uploads only
        rows [r * per, (r + 1) * per) over PCIe and the ranks all-gather the query rows over NVLink (one more collective,
        G times less H2D per rank), then per-shard search, packed all-gather, merge, D2H.  All device and pinned staging
        buffers are cached per shape, every copy is asynchronous on torch's current stream and the call synchronises once;
        a failure on any shard raises on every rank."""
        import torch
        import torch.distributed as dist

        dev = torch.device("cuda", torch.cuda.current_device())
        q = torch.from_numpy(queries)
        nq, dim = q.shape
        b = self._e2e_buffers(nq, dim, k, dev)
        if self.world > 1:
            per = -(-nq // self.world)
            lo, hi = min(self.rank * per, nq), min((self.rank + 1) * per, nq)
            mine = b["d_mine"]
            if hi > lo:
                mine[:hi - lo].copy_(q[lo:hi], non_blocking=True)
            dist.all_gather_into_tensor(b["d_q"], mine, group=self.group)
        else:
            b["d_q"][:nq].copy_(q, non_blocking=True)
        self.search_dev(b["d_q"][:nq], k, L_pq, b["d_i"], b["d_d"], k_shard=k_shard)
        b["h_i"].copy_(b["d_i"], non_blocking=True)
        b["h_d"].copy_(b["d_d"], non_blocking=True)
        err = None
        try:
            self.index.check(torch.cuda.current_stream())  # synchronises the stream, raises the reference's errors
        except rg.RgError as e:
            err = e
        raise_on_all_ranks(err, dev, self.group)
        return b["h_i"].numpy().view(np.uint32), b["h_d"].numpy()


class ReplicatedRoarIndex:
    """Query-sharded mode ("replicas"): every rank holds the FULL base and the same RoarGraph, owns a
    slice of the query batch, and the per-rank result blocks are collected with one all-gather — no merge step, and the
    ids are those of the single-GPU search.  The comparison point for the base-sharded contract: G x the memory, the
    same work per query, throughput ~ linear in G; base-sharding does G x the work per query at a smaller L_pq.

    Every rank must own the same number of queries (all_gather_into_tensor), so the batch is padded by repeating the
    last query of a short slice; search_dev returns only the real rows."""

    def __init__(self, index, group=None, local_search=None):
        self.index = index
        self.group = group
        self.world, self.rank = _world(group)
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
            return SEARCH_KERNELS
        per = -(-nq // self.world)
        lo = min(self.rank * per, nq)
        hi = min(lo + per, nq)
        key = (per, k, d_q_all.device)
        if key not in self._buf:
            self._buf[key] = (torch.empty((per, d_q_all.shape[1]), dtype=d_q_all.dtype, device=d_q_all.device),
                              torch.zeros((per, k), dtype=out_ids.dtype, device=d_q_all.device),
                              torch.zeros((per, k), dtype=out_dists.dtype, device=d_q_all.device),
                              torch.empty((per, k), dtype=torch.int64, device=d_q_all.device),
                              torch.empty((self.world * per, k), dtype=torch.int64, device=d_q_all.device))
        q_loc, loc_i, loc_d, packed, gathered = self._buf[key]
        launches = 0
        if hi - lo == per:
            self._search_local(d_q_all[lo:hi], k, L_pq, loc_i, loc_d)  # rows of a row-major batch: contiguous, in place
            launches = SEARCH_KERNELS
        elif hi > lo:
            q_loc[:hi - lo].copy_(d_q_all[lo:hi])
            q_loc[hi - lo:].copy_(d_q_all[hi - 1:hi].expand(per - (hi - lo), -1))
            self._search_local(q_loc, k, L_pq, loc_i, loc_d)
            launches = SEARCH_KERNELS
        pack_topk(loc_d, loc_i, 0, out=packed)
        dist.all_gather_into_tensor(gathered, packed, group=self.group)  # one collective for ids and distances together
        # rank r's block starts at r * per: contiguous in query order, padding only at the end
        w = gathered[:nq]
        out_ids.copy_((w & 0xffffffff).to(torch.int32))  # 0xffffffff wraps to the int32 view of the sentinel
        out_dists.copy_((w >> 32).to(torch.int32).view(torch.float32))
        return launches + 1


class ShardedKnn:
    """Brute-force kNN with the BASE sharded over the ranks of `group` (BASELINE.json configs[3]): every rank keeps its
    base rows [row_lo, row_lo + nb_local) prepared on its GPU (rg_knn_base_*: the parts of compute_groundtruth.cpp:396-416)
    and scores every query batch against them, ids offset by the part's start (:413); per batch ONE all-gather of the
    packed (distance, id)[batch][K] blocks and the k-way merge kernel (the std::sort merge of :424-444).

    slice_only: rank r merges (and returns) only queries [r * per, (r + 1) * per) of every batch — the layout of a
    10 M-query job whose result file is written in G pieces; otherwise every rank ends with the full answer.
    The exchange of batch b runs on a side stream while the tensor cores already score batch b + 1.

    `local_knn(base, q, K, metric, ids, dists, id_offset)` replaces the CUDA entry in the CPU tests."""

    def __init__(self, d_base_shard, metric, row_lo: int, group=None, local_knn=None, share_tau: bool = True,
                 resync: bool | None = None):
        import os
        self.group = group
        self.share_tau = share_tau
        self.world, self.rank = _world(group)
        self.metric = metric
        self.row_lo = int(row_lo)
        self._base = d_base_shard
        self._local = local_knn
        self._kb = None
        if local_knn is None:
            self._kb = rg.KnnBase(d_base_shard, metric)
        self._comm_stream = None
        # second agreement on the bound inside the shard (rg_knn_base_search_resync_dev): after a fifth of the SMALLEST
        # shard's rows — the same point on every rank, so the collectives pair up
        if resync is None:
            resync = os.environ.get("RG_KNN_RESYNC", "0") == "1"   # opt-in: ~2 % on one rank's share, not measured to pay at N = 8
        self.resync_rows = None
        if resync and share_tau and self.world > 1 and local_knn is None:
            import torch
            import torch.distributed as dist
            n = torch.tensor([d_base_shard.shape[0]], dtype=torch.int64, device=d_base_shard.device)
            dist.all_reduce(n, op=dist.ReduceOp.MIN, group=group)
            self.resync_rows = int(os.environ.get("RG_KNN_RESYNC_ROWS", int(n.item()) // 5))

    def close(self):
        if self._kb is not None:
            self._kb.close()
            self._kb = None

    def _start_bound(self, q, K, tau):
        """The shards agree on ONE float per query before they filter: each offers a bound of its ceil(K/G)-th best seed row,
        the maximum bounds the global K-th best distance (rg_knn_base_seed_dev); one all-reduce of nq floats, started
        asynchronously — search() issues it one batch AHEAD, so it travels under the previous batch's filter and is ordered
        before that batch's result exchange on the communicator."""
        import torch
        import torch.distributed as dist
        self._kb.seed_dev(q, K, -(-K // self.world), tau[:q.shape[0]], stream=torch.cuda.current_stream())
        return dist.all_reduce(tau[:q.shape[0]], op=dist.ReduceOp.MAX, group=self.group, async_op=True)

    def search(self, d_queries, K: int, batch: int = 151552, slice_only: bool = False, out=None):
        """-> (ids int32 [nq, K], dists fp32 [nq, K]) on the device: internal distances ascending (-IP for "ip", squared
        L2 for "l2"/"cosine"), global ids.  With slice_only the rows of the other ranks' slices are left untouched."""
        import torch
        import torch.distributed as dist

        nq = d_queries.shape[0]
        dev = d_queries.device
        G = self.world
        out_i, out_d = out if out is not None else (torch.empty((nq, K), dtype=torch.int32, device=dev),
                                                    torch.empty((nq, K), dtype=torch.float32, device=dev))
        cuda = d_queries.is_cuda
        if cuda and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
        shared = cuda and self._local is None and G > 1 and self.share_tau
        bufs = [None, None]
        events = [None, None]
        taus = [torch.empty(min(nq, batch), dtype=torch.float32, device=dev) for _ in range(2)] if shared else None
        bounds = [torch.empty(min(nq, batch), dtype=torch.float32, device=dev) for _ in range(2)] if shared else None
        starts = list(range(0, nq, batch))
        bound = self._start_bound(d_queries[0:min(nq, batch)], K, taus[0]) if shared and starts else None
        for bi, lo in enumerate(starts):
            hi = min(nq, lo + batch)
            n = hi - lo
            slot = bi & 1
            per = -(-n // G)
            if bufs[slot] is None:
                bufs[slot] = (torch.empty((batch, K), dtype=torch.int32, device=dev),
                              torch.empty((batch, K), dtype=torch.float32, device=dev),
                              torch.empty((G * (-(-batch // G)), K), dtype=torch.int64, device=dev),
                              torch.empty((G * (-(-batch // G)) if slice_only else G * batch, K), dtype=torch.int64, device=dev))
            loc_i, loc_d, packed, gathered = bufs[slot]
            if cuda and events[slot] is not None:
                torch.cuda.current_stream().wait_event(events[slot])  # the exchange that last used this slot has finished
            q = d_queries[lo:hi]
            if shared:
                bound.wait()                                         # this batch's bound (started one batch ago)
                tau = taus[slot]
                if bi + 1 < len(starts):
                    nlo = starts[bi + 1]
                    bound = self._start_bound(d_queries[nlo:min(nq, nlo + batch)], K, taus[slot ^ 1])
                if self.resync_rows is not None:
                    b = bounds[slot]
                    self._kb.search_resync_dev(
                        q, K, tau[:n], loc_i[:n], loc_d[:n], -(-K // G), self.resync_rows, b,
                        lambda off, cnt: dist.all_reduce(b[off:off + cnt], op=dist.ReduceOp.MAX, group=self.group),
                        id_offset=0, stream=torch.cuda.current_stream(), bound_rows=G * 65536)
                else:
                    self._kb.search_seeded_dev(q, K, tau[:n], loc_i[:n], loc_d[:n], id_offset=0,
                                               stream=torch.cuda.current_stream(), bound_rows=G * 65536)
            elif self._local is not None:
                self._local(self._base, q, K, self.metric, loc_i[:n], loc_d[:n], self.row_lo)
            else:
                self._kb.search_dev(q, K, loc_i[:n], loc_d[:n], id_offset=0, stream=torch.cuda.current_stream())
            if slice_only and G > 1 and G * per != n:
                packed[n:G * per].fill_(_PACKED_MISSING)             # ragged tail: the last ranks' slices are padded
            # the CPU checker already returns global ids
            pack_topk(loc_d[:n], loc_i[:n], 0 if self._local is not None else self.row_lo, out=packed[:n])

            def exchange():
                if slice_only and G > 1:
                    # a rank merges only its own slice of the batch: it needs that slice of every shard's block, nothing
                    # else — an all-to-all of [per, K] pieces, 1/G of the all-gather's traffic
                    g = gathered[:G * per]
                    dist.all_to_all_single(g, packed[:G * per], group=self.group)
                    q_lo, q_hi = min(self.rank * per, n), min((self.rank + 1) * per, n)
                    if q_hi > q_lo:
                        merge_packed(g.view(G, per, K), out_d[lo + q_lo:lo + q_lo + per], out_i[lo + q_lo:lo + q_lo + per], 0,
                                     q_hi - q_lo, stream=torch.cuda.current_stream() if cuda else None)
                    return
                g = gathered[:G * n]
                if G > 1:
                    dist.all_gather_into_tensor(g, packed[:n], group=self.group)
                else:
                    g.copy_(packed[:n])
                merge_packed(g.view(G, n, K), out_d[lo:hi], out_i[lo:hi], 0, n,
                             stream=torch.cuda.current_stream() if cuda else None)

            if cuda:
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream())
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(ready)
                    exchange()
                    events[slot] = torch.cuda.Event()
                    events[slot].record(self._comm_stream)
            else:
                exchange()
        if cuda:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        return out_i, out_d


def sharded_knn_dev(d_base_shard, d_queries, K: int, metric, row_lo: int, group=None, batch: int = 151552,
                    local_knn=None, slice_only: bool = False):
    """One-shot form of ShardedKnn (prepares the shard, searches, releases it)."""
    sk = ShardedKnn(d_base_shard, metric, row_lo, group=group, local_knn=local_knn)
    try:
        return sk.search(d_queries, K, batch=batch, slice_only=slice_only)
    finally:
        sk.close()


def query_sharded_knn_dev(d_base_full, d_queries, K: int, metric, group=None, local_knn=None):
    """Brute-force kNN with the QUERIES sharded and the base replicated — the comparison point SURVEY.md 8(e) names for
    the base-sharded contract above: rank r scores queries [r * per, (r + 1) * per) against ALL base rows and one
    all-gather collects the per-rank result blocks; no merge step, and every rank's tensor-core filter streams the whole
    base (the base-sharded form runs G times as many, G times smaller filter passes).  Needs the full base on every GPU.

    Returns (ids int32 [nq, K], dists fp32 [nq, K]) on the device, complete on every rank."""
    import torch
    import torch.distributed as dist

    world, rank = _world(group)
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
    packed = pack_topk(loc_d, loc_i, 0)
    gathered = torch.empty((world * per, K), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    w = gathered[:nq]  # rank r's block starts at r * per: query order, padding only behind the last query
    return (w & 0xffffffff).to(torch.int32), (w >> 32).to(torch.int32).view(torch.float32)
