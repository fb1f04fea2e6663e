#!/usr/bin/env python
"""One rank's share of the base-sharded kNN on ONE GPU: a 1/G shard of the base (default 1.25 M of 10 M rows), one
151 552-query batch, the shared bound emulated by this shard's own rank-ceil(K/G) seed value (the other shards' values are
statistically the same).  Times the seeded search under the chunk schedules (environment set by the caller).
usage: bench_knn_shard.py [SHARD_ROWS] [G]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import roargraph_b200 as rg
from roargraph_b200 import synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nq, K, dim = 151552, 100, 200
rg.init(0)
dev = torch.device("cuda", 0)
base, train, _ = synth.make_torch(rows * G, nq, 8, dim, dev)
shard = base[:rows].contiguous()
del base
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
kb = rg.KnnBase(shard, "ip")
ids = torch.empty((nq, K), dtype=torch.int32, device=dev); d = torch.empty((nq, K), dtype=torch.float32, device=dev)
tau = torch.empty(nq, dtype=torch.float32, device=dev)
out = {}
for name, bound_rows in (("growing", 0), ("whole_if_fits", G * 65536)):
    best = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        kb.seed_dev(train, K, -(-K // G), tau)
        kb.search_seeded_dev(train, K, tau, ids, d, bound_rows=bound_rows)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out[name] = {"ms": best * 1e3, "frac": 2.0 * nq * rows * dim / best / 1e12 / peaks["bf16_tflops_sustained"],
                 "short_lists": int((ids < 0).sum().item())}
# one re-sync inside the shard; the other shards' part of the maximum is emulated by this shard's own value
bound = torch.empty(nq, dtype=torch.float32, device=dev)
for frac in (0.1, 0.2):
    best = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        kb.seed_dev(train, K, -(-K // G), tau)
        kb.search_resync_dev(train, K, tau, ids, d, -(-K // G), int(rows * frac), bound, lambda off, cnt: None, bound_rows=0)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out[f"resync_at_{frac}"] = {"ms": best * 1e3, "frac": 2.0 * nq * rows * dim / best / 1e12 / peaks["bf16_tflops_sustained"]}
best = None
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    kb.search_dev(train, K, ids, d)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
out["local_tau_only"] = {"ms": best * 1e3, "frac": 2.0 * nq * rows * dim / best / 1e12 / peaks["bf16_tflops_sustained"]}
print(json.dumps({"rows": rows, "G": G, "env": {k: v for k, v in os.environ.items() if k.startswith("RG_KNN")}, **out}))
