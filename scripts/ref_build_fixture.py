#!/usr/bin/env python
"""Reference-side fixture at BASELINE configs[0] size (1M x 200, IP): data from the seeded numpy generator -> the
reference's compute_groundtruth (MKL) -> the reference's BuildRoarGraph CLI -> the reference's search CLI over the
README L_pq sweep.  Everything runs on the host cores with the UNMODIFIED reference binaries of oracle/_ref; nothing of
the product is involved.  Outputs:
  <out>/index.rg, base.fbin, test.fbin, test.gt, train.gt   (large; oracle/_ref/fixtures/ is git-ignored but travels to
                                                             the GPU box with the gpurun snapshot)
  tests/golden/ref_build_<tag>_stats.json                  (small, committed): degree histogram, ep, recall / cmps / hops
                                                            per L_pq, build wall time, core count
tests/test_gpu_build.py compares rg_index_build_roargraph on the same inputs against these statistics.

    python scripts/ref_build_fixture.py --nb 1000000 --tag 1M
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roargraph_b200 import formats as F, synth  # noqa: E402
from roargraph_b200.workload import README_L_SWEEP  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def run(cmd, **kw):
    t0 = time.time()
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)
    return time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--dist", default="ip")
    ap.add_argument("--tag", default="1M")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    out = a.out or os.path.join(REF, "fixtures", a.tag)
    os.makedirs(out, exist_ok=True)
    ntrain = a.nb // 10
    bf, tf, qf = (os.path.join(out, n) for n in ("base.fbin", "train.fbin", "test.fbin"))
    kf, gf, xf = (os.path.join(out, n) for n in ("train.gt", "test.gt", "index.rg"))
    timing = {}
    if not os.path.exists(bf):
        base, train, test = synth.make_numpy(a.nb, ntrain, a.nq, a.dim, seed=a.seed)
        F.write_fbin(bf, base)
        F.write_fbin(tf, train)
        F.write_fbin(qf, test)
        del base, train, test
    dist_fn = {"ip": "mips", "l2": "l2", "cosine": "cosine"}[a.dist]
    gt = os.path.join(REF, "ref_compute_groundtruth")
    if not os.path.exists(kf):
        timing["train_knn_s"] = run([gt, "--data_type", "float", "--dist_fn", dist_fn, "--base_file", bf, "--query_file", tf,
                                     "--gt_file", kf, "--K", "100"], stdout=subprocess.DEVNULL)
    if not os.path.exists(gf):
        timing["test_gt_s"] = run([gt, "--data_type", "float", "--dist_fn", dist_fn, "--base_file", bf, "--query_file", qf,
                                   "--gt_file", gf, "--K", "100"], stdout=subprocess.DEVNULL)
    if not os.path.exists(xf):
        log = os.path.join(out, "build.log")
        with open(log, "w") as lf:
            timing["build_s"] = run([os.path.join(REF, "ref_test_build_roargraph"), "--data_type", "float", "--dist", a.dist,
                                     "--base_data_path", bf, "--sampled_query_data_path", tf, "--projection_index_save_path", xf,
                                     "--learn_base_nn_path", kf, "--M_sq", "100", "--M_pjbp", "35", "--L_pjpq", "500",
                                     "-T", str(a.threads)], stdout=lf, stderr=subprocess.STDOUT)
    csv = os.path.join(out, "sweep.csv")
    for ln in (open(os.path.join(out, "build.log")) if os.path.exists(os.path.join(out, "build.log")) else []):
        if ln.startswith("indexing time:"):
            timing["build_indexing_s"] = float(ln.split(":")[1])
    if not os.path.exists(csv):
        timing["search_s"] = run([os.path.join(REF, "ref_test_search_roargraph"), "--data_type", "float", "--dist", a.dist,
                              "--base_data_path", bf, "--query_path", qf, "--gt_path", gf, "--projection_index_save_path", xf,
                              "--k", "10", "-T", str(a.threads), "--evaluation_save_path", csv, "--L_pq"]
                                 + [str(L) for L in README_L_SWEEP], stdout=subprocess.DEVNULL)
    ep, row_off, nbrs = F.read_projection_index(xf)
    deg = np.diff(row_off.astype(np.int64))
    rows = [ln.strip().split(",") for ln in open(csv) if ln.strip()]
    stats = {
        "what": "reference BuildRoarGraph + SearchRoarGraph (oracle/_ref, unmodified) on synth.make_numpy data",
        "nb": a.nb, "dim": a.dim, "ntrain": ntrain, "nq": a.nq, "dist": a.dist, "seed": a.seed, "threads": a.threads,
        "M_sq": 100, "M_pjbp": 35, "L_pjpq": 500, "ep": int(ep), "avg_degree": float(deg.mean()),
        "max_degree": int(deg.max()), "min_degree": int(deg.min()),
        "degree_hist": np.bincount(deg, minlength=int(deg.max()) + 1).tolist(),
        # CSV columns: L_pq,qps,avg_cmps,ms_per_query,recall,avg_hops  (tests/test_search_roargraph.cpp:234-235)
        "sweep": [{"L_pq": int(r[0]), "qps": float(r[1]), "avg_cmps": float(r[2]), "recall": float(r[4]),
                   "avg_hops": float(r[-1])} for r in rows],
        "timing": timing,
    }
    dst = os.path.join(ROOT, "tests", "golden", f"ref_build_{a.tag}_stats.json")
    with open(dst, "w") as f:
        json.dump(stats, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
