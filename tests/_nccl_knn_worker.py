"""Worker of tests/test_gpu_sharded_nccl.py (one process per GPU under torch.distributed.run): the base-sharded kNN over
NCCL — shared bound one batch ahead, packed exchange (all-gather, or all-to-all with slice_only), merge kernel — against the
single-GPU kNN over the whole base on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roargraph_b200 as rg  # noqa: E402
from roargraph_b200.sharded import ShardedKnn  # noqa: E402


def main():
    metric = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    rg.init(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nb, nq, dim, K, batch = 300_000, 10_001, 200, 100, 4096
    g = torch.Generator(device=dev)
    g.manual_seed(1234)                                   # the same data on every rank
    base = torch.randn((nb, dim), generator=g, device=dev)
    base *= 1.0 + 0.25 * torch.rand((nb, 1), generator=g, device=dev)
    q = torch.randn((nq, dim), generator=g, device=dev)
    per = -(-nb // world)
    lo, hi = rank * per, min(nb, (rank + 1) * per)
    shard = base[lo:hi].contiguous()
    want_i = torch.empty((nq, K), dtype=torch.int32, device=dev)
    want_d = torch.empty((nq, K), dtype=torch.float32, device=dev)
    rg.knn_dev(base, q, K, metric, want_i, want_d)
    ok = True
    for share, resync in ((True, True), (True, False), (False, False)):
        sk = ShardedKnn(shard, metric, lo, share_tau=share, resync=resync)
        got_i, got_d = sk.search(q, K, batch=batch)
        torch.cuda.synchronize()
        same = bool((got_i == want_i).all().item()) and bool((got_d == want_d).all().item())
        si = torch.full((nq, K), -7, dtype=torch.int32, device=dev)
        sd = torch.zeros((nq, K), dtype=torch.float32, device=dev)
        sk.search(q, K, batch=batch, slice_only=True, out=(si, sd))
        torch.cuda.synchronize()
        sk.close()
        mine = torch.zeros(nq, dtype=torch.bool, device=dev)
        for b0 in range(0, nq, batch):
            n = min(batch, nq - b0)
            p = -(-n // world)
            mine[b0 + min(rank * p, n):b0 + min((rank + 1) * p, n)] = True
        same_slice = bool((si[mine] == want_i[mine]).all().item()) and bool((sd[mine] == want_d[mine]).all().item())
        untouched = bool((si[~mine] == -7).all().item())
        flags = torch.tensor([int(same), int(same_slice), int(untouched)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"share_tau={share} resync={resync} full={int(flags[0])} slice={int(flags[1])} untouched={int(flags[2])}", flush=True)
        ok = ok and bool(flags.min().item())
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
