#!/usr/bin/env python
"""Kernel-only timing of the search kernel over launch-geometry knobs (RG_ROWS_PER_BATCH, RG_CTAS_PER_SM,
RG_WARPS_PER_CTA, RG_PREFETCH_ADJ) on one cached workload.  usage: tune_search.py NB [L ...]"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roargraph_b200 as rg
from roargraph_b200 import workload as W

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Ls = [int(x) for x in sys.argv[2:]] or [40, 200]
rg.init(0)
w = W.build_search_workload(nb, 200, "ip", int(os.environ.get("RG_TUNE_NTRAIN", max(1000, nb // 10))), 10000,
                            L_pjpq=int(os.environ.get("RG_TUNE_LPJPQ", 500)), cache_dir="/tmp/rg_bench_cache")
dev = torch.device("cuda", 0)
nq, k = 10000, 10
d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev); d_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
d_c = torch.empty(nq, dtype=torch.int32, device=dev); d_h = torch.empty(nq, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream()
ref = {}
print("L,R,ctas,warps,prefetch,visited,grid,smem,ms,qps,GBps,frac")
for L in Ls:
    for R, C, Wp, P, V in itertools.product((8, 16), (3, 4, 6), (4,), (0, 2), ("tags", "hash")):
        os.environ.update(RG_ROWS_PER_BATCH=str(R), RG_CTAS_PER_SM=str(C), RG_WARPS_PER_CTA=str(Wp), RG_PREFETCH_ADJ=str(P), RG_VISITED=V, RG_VTAB_MB="512")
        ts = []
        for rep in range(4):
            flush.fill_(rep)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); w.index.search_dev(w.d_queries, k, L, d_ids, d_d, d_c, d_h, stream=s); e1.record(s)
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ids = d_ids.cpu().numpy()
        if L not in ref: ref[L] = ids
        assert (ids == ref[L]).all(), "results changed with the launch geometry"
        st, ll = w.index.last_stats(), w.index.last_launch()
        ms = min(ts[1:]); gb = ((st["cmps"] + nq) * 800 + st["adj_words"] * 4) / ms / 1e6
        print(f"{L},{R},{C},{Wp},{P},{V},{ll['grid']},{ll['smem_bytes']},{ms:.3f},{nq / ms * 1e3:.0f},{gb:.0f},{gb / 6547.5:.3f}", flush=True)
