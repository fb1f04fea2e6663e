#!/usr/bin/env python
"""SearchBipartiteGraph at the size of BASELINE.json configs[1]: 10 M x 200 base, 1 M training queries,
raw bipartite graph of qbaseNNbipartite (src/index_bipartite.cpp:235-280, M_pjbp = 35), 10 000 test queries, top-10, ten
seeded entry points per query.  Prints one JSON line: the L_pq sweep (QPS through the host-buffer C-ABI call, recall@10,
avg cmps / hops, gather GB/s) — the paper's weak baseline beside RoarGraph's own numbers in profiles/r1_sweep_10M.csv.

    python scripts/bench_bipartite.py [--nb 10000000] [--out gpurun_out/bipartite_10M.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--M_pjbp", type=int, default=35)
    ap.add_argument("--ntrain", type=int, default=0, help="training queries (default nb / 10, README.md:77)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch

    import roargraph_b200 as rg
    from roargraph_b200 import synth
    from roargraph_b200.workload import bipartite_csr, recall_at_k, log

    rg.init(0)
    dev = torch.device("cuda", 0)
    ntrain, k = (a.ntrain or max(1000, a.nb // 10)), 10
    base, train, test = synth.make_torch(a.nb, ntrain, a.nq, a.dim, dev, seed=1234)
    gt_i = torch.empty((a.nq, k), dtype=torch.int32, device=dev)
    gt_d = torch.empty((a.nq, k), dtype=torch.float32, device=dev)
    rg.knn_dev(base, test, k, "ip", gt_i, gt_d)
    kn_i = torch.empty((ntrain, a.M_pjbp), dtype=torch.int32, device=dev)
    kn_d = torch.empty((ntrain, a.M_pjbp), dtype=torch.float32, device=dev)
    t0 = time.time()
    rg.knn_dev(base, train, a.M_pjbp, "ip", kn_i, kn_d)
    torch.cuda.synchronize()
    log(f"training-query kNN: {time.time() - t0:.1f}s")
    gt = gt_i.cpu().numpy().astype(np.uint32)
    t0 = time.time()
    row_off, nbrs = bipartite_csr(kn_i.cpu().numpy().view(np.uint32), a.nb, a.M_pjbp)
    log(f"bipartite graph on the host: {time.time() - t0:.1f}s, {len(nbrs)} edges, "
        f"{(np.diff(row_off[:a.nb + 1].astype(np.int64)) > 0).mean():.3f} of the base nodes have out-edges")
    del kn_i, kn_d, train
    rg.knn_release_workspace()
    idx = rg.RoarIndex(base, "ip")
    idx.set_bipartite(row_off, nbrs)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    nq = a.nq
    d_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_c = torch.empty(nq, dtype=torch.int32, device=dev)
    d_h = torch.empty(nq, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    rows = []
    for L in (10, 20, 50, 100, 200, 500, 1000, 2000):
        idx.search_bipartite_dev(test[:100], k, L, d_i[:100], d_d[:100], seed=7, stream=stream)  # the CLI's warm-up
        ts = []
        for rep in range(3):
            flush.fill_(rep)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            idx.search_bipartite_dev(test, k, L, d_i, d_d, d_c, d_h, seed=7, stream=stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 1e3)
        idx.check(stream)
        st = idx.last_stats()
        sec = min(ts)
        # algorithmic bytes: evaluated rows (incl. the 10 entry points) + the adjacency words walked
        alg = (st["cmps"] + nq * 10) * base.shape[1] * 4 + st["adj_words"] * 4
        rows.append({"L_pq": L, "qps": nq / sec, "recall": recall_at_k(d_i.cpu().numpy().view(np.uint32), gt, k),
                     "avg_cmps": st["cmps"] / nq, "avg_hops": st["hops"] / nq, "kernel_ms": sec * 1e3,
                     "gather_GBps": alg / sec / 1e9, "hbm_frac": alg / sec / 1e9 / peak})
        log(str(rows[-1]))
    line = {"path": "SearchBipartiteGraph (A2)", "workload": f"{a.nb}x{a.dim} ip, {ntrain} training queries, M_pjbp={a.M_pjbp}, "
            f"{a.nq} queries top-{k}, 10 seeded entry points", "timing": "CUDA events around rg_search_bipartite_dev, "
            "256 MiB L2 flush between repetitions, best of 3", "sweep": rows}
    print(json.dumps(line), flush=True)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
