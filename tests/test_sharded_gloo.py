"""CPU, world_size 2 over gloo: the host-side plumbing of the base-sharded search (shard ranges, local->global ids,
all-gather layout) with per-shard results produced by the ORACLE on each rank's slice; the merged answer (numpy
checker in this test) must equal brute force over the union of the shards' candidates and beat single-shard recall."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GoldenCase
from roargraph_b200.sharded import gather_topk, shard_range, sharded_knn_dev


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    O = Oracle()
    g = GoldenCase("ip_d40")
    nb = g.base.shape[0]
    lo, hi = shard_range(nb, rank, world)
    k = 10
    # each rank: exact top-k of its slice (stand-in for the per-shard index search), ids made global
    ids, dists = O.exact_knn(g.base[lo:hi], g.queries, k, "ip")
    loc_i = torch.from_numpy(ids.astype(np.int32) + lo)
    loc_d = torch.from_numpy(dists)
    all_d, all_i = gather_topk(loc_d, loc_i)
    assert all_d.shape == (world, g.queries.shape[0], k)
    # every rank sees rank r's block at index r
    assert torch.equal(all_i[rank], loc_i) and torch.equal(all_d[rank], loc_d)
    if rank == 0:
        np.savez(out, d=all_d.numpy(), i=all_i.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_rows():
    for nb, G in ((10, 3), (10_000_000, 8), (7, 8)):
        r = [shard_range(nb, s, G) for s in range(G)]
        assert r[0][0] == 0 and r[-1][1] == nb
        assert all(r[i][1] == r[i + 1][0] for i in range(G - 1))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_gather_and_merge(tmp_path, oracle):
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    g = GoldenCase("ip_d40")
    G, nq, k = z["d"].shape
    alld = z["d"].transpose(1, 0, 2).reshape(nq, G * k)
    alli = z["i"].transpose(1, 0, 2).reshape(nq, G * k)
    order = np.lexsort((alli, alld), axis=1)[:, :k]  # (distance asc, id asc), as rg_topk_merge_dev
    merged_i = np.take_along_axis(alli, order, 1)
    merged_d = np.take_along_axis(alld, order, 1)
    full_i, full_d = oracle.exact_knn(g.base, g.queries, k, "ip")
    np.testing.assert_allclose(merged_d, full_d, rtol=1e-6, atol=1e-6)
    assert (merged_i == full_i).mean() > 0.999


def _knn_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    O = Oracle()
    g = GoldenCase("l2_d104")
    nb = g.base_raw.shape[0]
    lo, hi = shard_range(nb, rank, world)
    K = 7

    def local_knn(b, q, k, m, ids, dists, id_offset):  # checker standing in for rg.knn_dev on this rank's part
        i, d = O.exact_knn(b.numpy(), q.numpy(), k, m)
        ids.copy_(torch.from_numpy(i.astype(np.int32) + id_offset))
        dists.copy_(torch.from_numpy(d))

    ids, dists = sharded_knn_dev(torch.from_numpy(g.base_raw[lo:hi]), torch.from_numpy(g.test_raw), K, "l2", lo, batch=17,
                                 local_knn=local_knn)
    # slice_only: every rank merges only its share of each 17-query batch; the union over the ranks is the full answer
    si, sd = sharded_knn_dev(torch.from_numpy(g.base_raw[lo:hi]), torch.from_numpy(g.test_raw), K, "l2", lo, batch=17,
                             local_knn=local_knn, slice_only=True)
    nq = g.test_raw.shape[0]
    mine = np.zeros(nq, dtype=bool)
    for b0 in range(0, nq, 17):
        n = min(17, nq - b0)
        per = -(-n // world)
        mine[b0 + min(rank * per, n):b0 + min((rank + 1) * per, n)] = True
    assert torch.equal(si[torch.from_numpy(mine)], ids[torch.from_numpy(mine)])
    assert torch.equal(sd[torch.from_numpy(mine)], dists[torch.from_numpy(mine)])
    cover = torch.from_numpy(mine.astype(np.int32))
    dist.all_reduce(cover)
    assert bool((cover == 1).all())  # every query belongs to exactly one rank's slice
    if rank == 1:  # every rank holds the full answer
        np.savez(out, d=dists.numpy(), i=ids.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_knn(tmp_path, oracle):
    """Base parts on two ranks, ragged query batches (17 per all-gather): equals the single-part exact kNN."""
    out = str(tmp_path / "knn.npz")
    mp.spawn(_knn_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    g = GoldenCase("l2_d104")
    full_i, full_d = oracle.exact_knn(g.base_raw, g.test_raw, 7, "l2")
    np.testing.assert_allclose(z["d"], full_d, rtol=1e-5, atol=1e-5)
    assert (z["i"] == full_i).mean() > 0.999


def test_three_rank_sharded_knn(tmp_path, oracle):
    """Three ranks: 17-query batches do not divide by the world size, so the all-to-all slices are padded (sentinel words) and
    the last rank's slice is short; same answer."""
    out = str(tmp_path / "knn3.npz")
    mp.spawn(_knn_worker, args=(3, _free_port(), out), nprocs=3, join=True)
    z = np.load(out)
    g = GoldenCase("l2_d104")
    full_i, full_d = oracle.exact_knn(g.base_raw, g.test_raw, 7, "l2")
    np.testing.assert_allclose(z["d"], full_d, rtol=1e-5, atol=1e-5)
    assert (z["i"] == full_i).mean() > 0.999


def _replica_worker(rank, world, port, out, nq):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from roargraph_b200.sharded import ReplicatedRoarIndex
    O = Oracle()
    g = GoldenCase("ip_d40")
    k, L = 10, 32
    calls = []

    def local_search(d_q, k_, L_, out_ids, out_dists):  # checker standing in for the CUDA search on this rank's slice
        r = O.search_roargraph(g.base, g.row_off, g.nbrs, g.ep, g.metric, d_q.numpy(), k_, L_)
        out_ids.copy_(torch.from_numpy(r["ids"].astype(np.int32)))
        out_dists.copy_(torch.from_numpy(r["dists"]))
        calls.append(d_q.shape[0])

    ridx = ReplicatedRoarIndex(None, local_search=local_search)
    q = torch.from_numpy(g.queries[:nq])
    ids = torch.empty((nq, k), dtype=torch.int32)
    dists = torch.empty((nq, k), dtype=torch.float32)
    ridx.search_dev(q, k, L, ids, dists)
    per = -(-nq // world)
    assert calls == ([per] if rank * per < nq else [])  # every rank searches only its (padded) slice
    if rank == 1:  # every rank holds the complete answer
        np.savez(out, i=ids.numpy(), d=dists.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [64, 33, 1])
def test_two_rank_query_sharded_replicas(tmp_path, oracle, nq):
    """Replica mode: queries split over two ranks (ragged: 33 -> 17 + 16, 1 -> 1 + 0), one all-gather, no merge;
    the result equals the single-process search id for id."""
    out = str(tmp_path / "replica.npz")
    mp.spawn(_replica_worker, args=(2, _free_port(), out, nq), nprocs=2, join=True)
    z = np.load(out)
    g = GoldenCase("ip_d40")
    want = oracle.search_roargraph(g.base, g.row_off, g.nbrs, g.ep, g.metric, g.queries[:nq], 10, 32)
    assert (z["i"].view(np.uint32) == want["ids"]).all()
    assert (z["d"].view(np.uint32) == want["dists"].view(np.uint32)).all()


def _qknn_worker(rank, world, port, out, nq):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from roargraph_b200.sharded import query_sharded_knn_dev
    O = Oracle()
    g = GoldenCase("l2_d104")
    K = 9

    def local_knn(b, q, k, m, ids, dists, id_offset):  # checker standing in for rg.knn_dev
        i, d = O.exact_knn(b.numpy(), q.numpy(), k, m)
        ids.copy_(torch.from_numpy(i.astype(np.int32) + id_offset))
        dists.copy_(torch.from_numpy(d))

    ids, dists = query_sharded_knn_dev(torch.from_numpy(g.base_raw), torch.from_numpy(g.test_raw[:nq]), K, "l2",
                                       local_knn=local_knn)
    assert ids.shape == (nq, K)
    if rank == 0:
        np.savez(out, d=dists.numpy(), i=ids.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [48, 35, 1])
def test_two_rank_query_sharded_knn(tmp_path, oracle, nq):
    """Queries split over two ranks (ragged, and a rank with nothing to do), base replicated: equals the one-process kNN."""
    out = str(tmp_path / "qknn.npz")
    mp.spawn(_qknn_worker, args=(2, _free_port(), out, nq), nprocs=2, join=True)
    z = np.load(out)
    g = GoldenCase("l2_d104")
    full_i, full_d = oracle.exact_knn(g.base_raw, g.test_raw[:nq], 9, "l2")
    assert (z["i"] == full_i).all() and (z["d"].view(np.uint32) == full_d.view(np.uint32)).all()


def _sharded_search_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from roargraph_b200 import RgError
    from roargraph_b200.sharded import ShardedRoarIndex, raise_on_all_ranks
    O = Oracle()
    g = GoldenCase("ip_d40")
    nb = g.base.shape[0]
    lo, hi = shard_range(nb, rank, world)
    nq = g.queries.shape[0]

    def local_search(d_q, k, L, out_ids, out_dists):  # exact top-k of the slice stands in for the per-shard index search;
        ids, dists = O.exact_knn(g.base[lo:hi], d_q.numpy(), k, "ip")  # rank 1 pretends query 3 found only 2 results
        ids = ids.astype(np.int32)
        if rank == 1:
            ids[3, 2:] = -1
            dists[3, 2:] = np.inf
        out_ids.copy_(torch.from_numpy(ids))
        out_dists.copy_(torch.from_numpy(dists))

    sidx = ShardedRoarIndex(None, row_lo=lo, local_search=local_search)
    res = {}
    for ks in (10, 6):
        ids = torch.empty((nq, 10), dtype=torch.int32)
        dists = torch.empty((nq, 10), dtype=torch.float32)
        assert sidx.search_dev(torch.from_numpy(g.queries), 10, 32, ids, dists, k_shard=ks) == 5  # tag maintenance (2) + search + pack + merge
        res[ks] = (ids.numpy().copy(), dists.numpy().copy())
    if rank == 0:
        np.savez(out, i10=res[10][0], d10=res[10][1], i6=res[6][0], d6=res[6][1])
    # a failure on one rank raises on every rank instead of leaving the peers in the next collective
    err = RgError(-3, "not enough results: 2, expected: 10") if rank == 1 else None
    try:
        raise_on_all_ranks(err, torch.device("cpu"))
        raised = False
    except RgError:
        raised = True
    assert raised
    raise_on_all_ranks(None, torch.device("cpu"))  # and nothing happens when every rank is fine
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_packed_exchange(tmp_path, oracle):
    """ShardedRoarIndex over gloo: ONE packed all-gather, global ids, the short-result sentinel is neither offset nor
    merged in front of real results, k_shard < k keeps the k best of 2 x k_shard, errors surface on every rank."""
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_sharded_search_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    g = GoldenCase("ip_d40")
    nb = g.base.shape[0]
    full_i, full_d = oracle.exact_knn(g.base, g.queries, 10, "ip")
    keep = np.ones(len(full_i), dtype=bool)
    keep[3] = False  # the query rank 1 truncated
    np.testing.assert_allclose(z["d10"][keep], full_d[keep], rtol=1e-6, atol=1e-6)
    assert (z["i10"][keep] == full_i[keep]).mean() > 0.999
    assert (z["i10"] != -1).all() and (z["i10"] >= 0).all() and (z["i10"] < nb).all()  # 2 x 10 entries: the sentinel never surfaces
    assert np.isfinite(z["d10"]).all() and (np.diff(z["d10"], axis=1) >= 0).all()
    # k_shard = 6: the 10 best of 2 x 6; equal to the exact answer wherever neither shard holds more than 6 of the top-10
    lo1 = nb // 2
    share = (full_i < lo1).sum(1)
    ok = keep & (share <= 6) & (share >= 4)
    assert ok.sum() > 10 and (z["i6"][ok] == full_i[ok]).mean() > 0.999
