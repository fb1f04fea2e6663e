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
from roargraph_b200.sharded import gather_topk, shard_range


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
