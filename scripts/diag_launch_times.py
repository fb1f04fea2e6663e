#!/usr/bin/env python
"""Per-launch device times of the search kernel under different host patterns between launches (back-to-back, a device
sync, a sync plus an idle gap, an L2 flush): looks for what makes single launches slow.
usage: diag_launch_times.py NB L N"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roargraph_b200 as rg
from roargraph_b200 import workload as W
nb = int(sys.argv[1]); L = int(sys.argv[2]); N = int(sys.argv[3])
rg.init(0)
w = W.build_search_workload(nb, 200, "ip", max(1000, nb // 10), 10000, cache_dir="/tmp/rg_bench_cache")
dev = torch.device("cuda", 0)
nq, k = 10000, 10
d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev); d_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream()

def run(name, sync, gap_ms, do_flush):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
    for i in range(N):
        if do_flush:
            flush.fill_(i & 0xff)
        if sync:
            torch.cuda.synchronize()
        if gap_ms:
            time.sleep(gap_ms / 1e3)
        evs[i][0].record(s); w.index.search_dev(w.d_queries, k, L, d_ids, d_d, stream=s); evs[i][1].record(s)
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in evs])[3:]
    print(f"{name:34s} median {np.median(t):7.3f} mean {t.mean():7.3f} p90 {np.percentile(t, 90):7.3f} max {t.max():7.3f}  >1.15x median: {(t > 1.15 * np.median(t)).sum()}/{len(t)}", flush=True)

run("back-to-back", False, 0, False)
run("flush, back-to-back", False, 0, True)
run("sync each", True, 0, False)
run("flush + sync each", True, 0, True)
run("sync + 1 ms idle", True, 1, False)
run("sync + 10 ms idle", True, 10, False)
run("flush + sync + 10 ms idle", True, 10, True)
run("back-to-back again", False, 0, False)
