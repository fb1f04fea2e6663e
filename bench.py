#!/usr/bin/env python
"""bench.py — QPS of the batched RoarGraph beam search at recall@10 >= 0.9 on t2i-10M-shaped synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nb 10000000 --dim 200 --dist ip ...]

One "step" = one pass of SearchRoarGraph over the whole 10 000-query batch at the smallest L_pq of the reference's
canonical sweep (README.md:118) whose recall@10 against exact ground truth is >= 0.9.
  value     queries/s with queries and outputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same through the host-buffer entry point rg_search_roargraph (C-ABI): pinned host queries -> H2D ->
            kernel -> D2H of ids/dists/cmps/hops, wall clock around the synchronous call
  roofline  algorithmic gather bytes of the search kernel / its CUDA-event duration vs the measured HBM copy peak
  cpu_baseline  the reference's own SearchRoarGraph (compiled unmodified, oracle/_ref) on the host cores, same index
--impl reference prints the reference-CPU arm as the headline line instead (rank 0 only under torchrun).
N > 1: base rows sharded over the ranks, per-shard index, one NCCL all-gather + merge kernel per step (strong scaling);
--shard query runs the comparison point instead: the full index on every rank, --nq queries per rank, one all-gather.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nb", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--dist", default="ip", choices=["ip", "l2", "cosine"])
    ap.add_argument("--ntrain", type=int, default=0, help="training queries (default nb/10, README.md:77)")
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--M_sq", type=int, default=100)
    ap.add_argument("--M_pjbp", type=int, default=35)
    ap.add_argument("--L_pjpq", type=int, default=500)
    ap.add_argument("--target_recall", type=float, default=0.9)
    ap.add_argument("--L_pq", type=int, default=0, help="fix L_pq instead of calibrating it on recall")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cache_dir", default=os.environ.get("RG_BENCH_CACHE", "/tmp/rg_bench_cache"))
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_knn", action="store_true", help="skip the brute-force kNN leg")
    ap.add_argument("--shard", default="base", choices=["base", "query"],
                    help="N > 1: 'base' = base rows sharded, per-shard index, all-gather + merge (the contract, strong "
                         "scaling); 'query' = full index replicated, --nq queries PER RANK, all-gather only (weak scaling)")
    ap.add_argument("--sweep", default="", help="also write the full L_pq sweep table (CSV) to this path")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, device_index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.idx, self.period = device_index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remapping: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
                        break
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML sampling unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(a):
    label = "t2i" if (a.dim == 200 and a.dist == "ip") else ("laion" if a.dim == 512 else "synthetic")
    return (f"{label}-like synthetic {a.nb}x{a.dim} {a.dist}, {a.nq} queries top-{a.k}, RoarGraph M_sq={a.M_sq} "
            f"M_pjbp={a.M_pjbp} L_pjpq={a.L_pjpq} from {a.ntrain} training queries")


def load_tensor_peak(burst=False):
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            key = "bf16_tflops" if burst else "bf16_tflops_sustained"
            return float(json.load(f)[key]), f"measured (MEASURED_PEAKS.json {key})"
    return (1650.0 if burst else 1400.0), "fallback (B200_PROFILING.md, " + ("burst)" if burst else "sustained)")


def time_knn(a, w, dev, world, K=100, reps=2, query_sharded=False):
    """BASELINE.json's second metric: brute-force training-query x base kNN pairs/s (compute_groundtruth's exact_knn),
    base sharded over the ranks like the search (configs[3]): local tcgen05 filter + exact re-rank, all-gather, merge."""
    import torch
    import torch.distributed as dist

    from roargraph_b200.sharded import query_sharded_knn_dev, sharded_knn_dev

    q = w.d_train_sample
    metric = "l2" if a.dist == "l2" else "ip"
    best = None
    for _ in range(reps + 1):  # first pass = warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        if query_sharded:  # replicas: every rank holds the whole base and scores its slice of the queries
            ids, dists = query_sharded_knn_dev(w.d_base, q, K, metric)
        else:
            ids, dists = sharded_knn_dev(w.d_base, q, K, metric, w.row_lo)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        best = dt if best is None else min(best, dt)
        del ids, dists
    pairs = float(q.shape[0]) * float(w.nd_total)
    peak, src = load_tensor_peak()
    burst, _ = load_tensor_peak(burst=True)
    tf = 2.0 * pairs * a.dim / best / 1e12
    return {"metric": "brute-force kNN pairs/s", "value": pairs / best, "unit": "pairs/s", "nq": int(q.shape[0]),
            "nb": int(w.nd_total), "dim": a.dim, "K": K, "dist": metric, "seconds": best, "tflops": tf,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak * world, "unit": "TFLOP/s",
                         "frac": tf / (peak * world), "peak_source": src,
                         "peak_burst": burst * world, "frac_of_burst": tf / (burst * world),
                         "kernel": "knn_tc_filter_pair_kernel (tcgen05 bf16 filter) + knn_merge_fast_kernel (fp32 re-rank)"},
            "note": "wall clock of the whole call: fp32->bf16 conversion, norms, every filter and merge launch"
                    + ((", queries split over the ranks, all-gather of the result blocks" if query_sharded else
                        ", all-gather + k-way merge over the shards") if world > 1 else "")}


def time_reference_cpu(a, w, L, steps, warmup, budget_s):
    """The unmodified reference SearchRoarGraph (oracle/_ref/libref_roargraph.so) on the host cores over this index, under
    the reference CLI's loop shape (100-query serial warm-up, then omp parallel for schedule(dynamic,1),
    tests/test_search_roargraph.cpp:198-209).  One step = one pass over a bounded sample of the query batch (the first
    n_sample queries, sized so that warmup + steps passes take about budget_s); `warmup` untimed passes, then exactly
    `steps` timed ones; value = n_sample * steps / their total wall time."""
    from oracle.oracle import Reference
    from roargraph_b200 import formats as F
    from roargraph_b200.workload import log, recall_at_k

    R = Reference()
    cores = R.num_procs()
    t0 = time.time()
    ep, row_off, nbrs = w.index.get_graph()
    os.makedirs(a.cache_dir, exist_ok=True)
    ipath = os.path.join(a.cache_dir, f"ref_index_{os.getpid()}.rg")
    F.write_projection_index(ipath, ep, row_off, nbrs)
    del row_off, nbrs
    base_host = w.d_base.cpu().numpy()
    h = R.index_from_memory(base_host, ipath, "l2" if a.dist == "l2" else "ip", cores)
    os.remove(ipath)
    log(f"reference index in host memory ({time.time() - t0:.1f}s), {cores} host threads")
    q = w.queries_host
    # size the sample on a small slice (this call also does the CLI's 100-query serial warm-up)
    r0 = R.search_roargraph(h, q[:min(len(q), 1000)], a.k, L, cores, warmup=True)
    qps_est = min(len(q), 1000) / max(r0["wall_s"], 1e-6)
    n_sample = int(min(len(q), max(1000, qps_est * budget_s / max(1, steps + warmup))))
    for _ in range(warmup):
        R.search_roargraph(h, q[:n_sample], a.k, L, cores, warmup=False)
    walls, res = [], None
    for _ in range(steps):
        res = R.search_roargraph(h, q[:n_sample], a.k, L, cores, warmup=False)
        walls.append(res["wall_s"])
    R.index_free(h)
    total = float(sum(walls))
    rec = recall_at_k(res["ids"], w.gt_ids[:n_sample], a.k)
    return {"value": n_sample * steps / total, "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": f"first {n_sample} of {len(q)} queries at L_pq={L} per step, {steps} timed passes after {warmup} "
                      f"untimed ones and the CLI's 100-query serial warm-up; recall@{a.k}={rec:.4f}",
            "ms_per_step": total / steps * 1e3, "step_ms": [round(x * 1e3, 3) for x in walls], "n_sample": n_sample,
            "res": res}


def main():
    a = parse_args()
    if a.ntrain <= 0:
        a.ntrain = max(1000, a.nb // 10)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 alone

    import torch
    import torch.distributed as dist

    import roargraph_b200 as rg
    from roargraph_b200 import workload as W
    from roargraph_b200.sharded import ReplicatedRoarIndex, ShardedRoarIndex

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    rg.init(local_rank)
    dev = torch.device("cuda", local_rank)
    n_shards = 1
    if world > 1 and a.impl == "b200":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        n_shards = world

    replicas = world > 1 and a.impl == "b200" and a.shard == "query"
    if replicas:  # full index on every rank (same seed -> same data, deterministic build -> same graph), queries split
        n_shards = 1
    nq = a.nq * (world if replicas else 1)
    w = W.build_search_workload(a.nb, a.dim, a.dist, a.ntrain, nq, a.k, a.M_sq, a.M_pjbp, a.L_pjpq, a.seed, local_rank,
                                shard=rank if n_shards > 1 else 0, n_shards=n_shards, cache_dir=a.cache_dir)
    k = a.k
    d_q = w.d_queries
    if replicas:
        ridx = ReplicatedRoarIndex(w.index)
        d_q_local = d_q[rank * a.nq:(rank + 1) * a.nq]  # this rank's rows (a view); also the kernel-only / roofline pass
    else:
        sidx = ShardedRoarIndex(w.index, w.row_lo)
        d_q_local = d_q
    nq_local = d_q_local.shape[0]
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_dists = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_cmps = torch.empty(nq, dtype=torch.int32, device=dev)
    d_hops = torch.empty(nq, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def run_dev(L):
        if replicas:
            return ridx.search_dev(d_q, k, L, d_ids, d_dists) + 1  # + the all-gather
        return sidx.search_dev(d_q, k, L, d_ids, d_dists, d_cmps, d_hops)

    # ---- calibrate L_pq on recall (outside the timed region) ------------------------------------------------------
    sweep_rows = []
    L_star, rec_star = a.L_pq, None
    if a.L_pq <= 0:
        for L in W.README_L_SWEEP:
            if L < k:
                continue
            run_dev(L)
            w.index.check(stream)
            rec = W.recall_at_k(d_ids.cpu().numpy().view(np.uint32), w.gt_ids, k)
            sweep_rows.append((L, rec))
            if rec >= a.target_recall:
                L_star, rec_star = L, rec
                break
        if rec_star is None:
            L_star, rec_star = sweep_rows[-1]
            W.log(f"WARNING: recall {a.target_recall} not reached inside the sweep; using L_pq={L_star} (recall {rec_star:.4f})")
    else:
        run_dev(L_star)
        w.index.check(stream)
        rec_star = W.recall_at_k(d_ids.cpu().numpy().view(np.uint32), w.gt_ids, k)
    W.log(f"L_pq={L_star}: recall@{k}={rec_star:.4f}")
    L = L_star

    if a.impl == "reference":
        cpu = time_reference_cpu(a, w, L, steps=a.steps, warmup=a.warmup, budget_s=90.0)
        line = {
            "impl": "reference", "metric": f"QPS@recall{k}>={a.target_recall}", "value": cpu["value"], "unit": "queries/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": cpu["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "L_pq": L, "recall": rec_star,
                       "note": "unmodified reference SearchRoarGraph on the host cores; inputs (base, index, ground "
                               "truth) were produced on the GPU before the timed region"},
            "cpu_baseline": {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cpu["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "step_ms": cpu["step_ms"],
        }
        print(json.dumps(line), flush=True)
        return 0

    # ---- timed region: device-resident -----------------------------------------------------------------------------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    for _ in range(max(a.warmup, 3)):
        run_dev(L)
    w.index.check(stream)
    sampler = ClockSampler(local_rank)
    if sampler.ok:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches = 0
    for s in range(a.steps):
        flush.fill_(s & 0xff)  # evict the previous step's rows from L2 (not timed)
        ev[s][0].record(stream)
        launches += run_dev(L)
        ev[s][1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    w.index.check(stream)
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / a.steps
    qps = nq / (ms_per_step / 1e3)
    # kernel-only duration for the roofline (search kernel alone on this rank, same stream, L2 flushed)
    kern_ms = []
    loc_i = torch.empty((nq_local, k), dtype=torch.int32, device=dev)
    loc_d = torch.empty((nq_local, k), dtype=torch.float32, device=dev)
    for s in range(min(a.steps, 10)):
        flush.fill_(s & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        w.index.search_dev(d_q_local, k, L, loc_i, loc_d, d_cmps[:nq_local], d_hops[:nq_local], stream=stream)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        kern_ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler.ok else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                                                "note": getattr(sampler, "err", "NVML unavailable")}
    # median: every 255th query of a slot clears its visited-tag array (VisitedList::reset), which lands in one launch
    kern_avg_ms = float(np.median(kern_ms))
    st = w.index.last_stats()   # of the kernel-only pass: this rank's queries on this rank's index
    ll = w.index.last_launch()
    alg_bytes = (st["cmps"] + nq_local) * w.d_base.shape[1] * 4 + st["adj_words"] * 4
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (kern_avg_ms / 1e3) / 1e9

    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of this very workload, if there is one
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{a.nb}x{a.dim}:{a.dist}:L{L}", {}).get("traffic_bytes") if world == 1 else None
    except OSError:
        pass

    # ---- end to end through the C-ABI host-buffer call ---------------------------------------------------------------
    hq = torch.from_numpy(w.queries_host).pin_memory()
    h_ids = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    h_d = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    h_c = torch.empty(nq, dtype=torch.int32).pin_memory()
    h_h = torch.empty(nq, dtype=torch.int32).pin_memory()
    e2e_ms = []
    if world == 1:
        for s in range(a.warmup + a.steps):
            flush.fill_(s & 0xff)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            w.index.search_into(hq.numpy(), k, L, h_ids.numpy(), h_d.numpy(), h_c.numpy(), h_h.numpy())
            t1 = time.perf_counter()
            if s >= a.warmup:
                e2e_ms.append((t1 - t0) * 1e3)
        h2d = hq.numel() * 4
        d2h = (h_ids.numel() + h_d.numel() + h_c.numel() + h_h.numel()) * 4
        e2e_ids = h_ids.numpy().view(np.uint32).copy()
    else:
        for s in range(a.warmup + a.steps):
            flush.fill_(s & 0xff)
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            if replicas:  # this rank's slice H2D, search, all-gather, the whole result D2H
                d_q_local.copy_(hq[rank * a.nq:(rank + 1) * a.nq], non_blocking=True)
                ridx.search_dev(d_q, k, L, d_ids, d_dists)
                h_ids.copy_(d_ids, non_blocking=True)
                h_d.copy_(d_dists, non_blocking=True)
                w.index.check(stream)
                e2e_ids = h_ids.numpy().view(np.uint32)
            else:
                e2e_ids, _ = sidx.search(hq.numpy(), k, L)
            t1 = time.perf_counter()
            if s >= a.warmup:
                e2e_ms.append((t1 - t0) * 1e3)
        t = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = [float(t.item())]
        h2d = hq.numel() * 4 // (world if replicas else 1)
        d2h = nq * k * 8
    e2e_qps = nq / (float(np.mean(e2e_ms)) / 1e3)
    e2e_recall = W.recall_at_k(e2e_ids, w.gt_ids, k)

    sweep_table = None
    if a.sweep and rank == 0 and world == 1:
        rows = []
        for Ls in W.README_L_SWEEP:
            if Ls < k:
                continue
            ts = []
            for rep in range(3):
                flush.fill_(rep)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                w.index.search_dev(d_q, k, Ls, d_ids, d_dists, d_cmps, d_hops, stream=stream)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1))
            s2 = w.index.last_stats()
            rec = W.recall_at_k(d_ids.cpu().numpy().view(np.uint32), w.gt_ids, k)
            ms = min(ts)
            gbs = ((s2["cmps"] + nq) * w.d_base.shape[1] * 4 + s2["adj_words"] * 4) / (ms / 1e3) / 1e9
            rows.append((Ls, nq / (ms / 1e3), s2["cmps"] / nq, ms / nq, rec, s2["hops"] / nq, gbs, gbs / peak))
        os.makedirs(os.path.dirname(os.path.abspath(a.sweep)), exist_ok=True)
        with open(a.sweep, "w") as f:
            f.write("L_pq,qps,avg_cmps,ms_per_query,recall,avg_hops,gather_GBps,hbm_frac\n")  # CLI columns + roofline
            for r in rows:
                f.write(",".join(f"{x:.6g}" for x in r) + "\n")
        sweep_table = a.sweep

    knn_out = None
    if not a.no_knn:  # base-sharded like the search (configs[3]); with --shard query the queries are split instead
        try:
            knn_out = time_knn(a, w, dev, world, query_sharded=replicas)
        except Exception as e:  # noqa: BLE001
            knn_out = {"error": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            cpu = time_reference_cpu(a, w, L, steps=5, warmup=1, budget_s=20.0)
            n_s = cpu["n_sample"]
            same = bool((cpu["res"]["ids"] == e2e_ids[:n_s]).all())
            cpu_out = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            cpu_out["ids_identical_to_gpu"] = same
        except Exception as e:  # noqa: BLE001
            cpu_out = {"value": None, "unit": "queries/s", "cores": None, "kind": "reference", "sample": f"failed: {e!r}"}
    else:
        cpu_out = None

    if rank == 0:
        line = {
            "metric": f"QPS@recall{k}>={a.target_recall}", "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if replicas else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "L_pq": L, "recall": rec_star, "e2e_recall": e2e_recall,
                       "sharding": (f"queries over {world} replicas of the full index ({a.nq} queries per rank), all-gather"
                                    if replicas else
                                    f"base rows over {world} rank(s), per-shard index, all-gather + merge" if world > 1
                                    else "none"),
                       "l2": "256 MiB fill between steps (outside the timed events); touched rows >> 126 MB L2",
                       "avg_cmps": st["cmps"] / nq_local, "avg_hops": st["hops"] / nq_local, "launch": ll, "setup": w.info},
            "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": float(np.mean(e2e_ms)), "step_ms": [round(float(x), 3) for x in e2e_ms]},
            "gpu_launches": int(launches),
            "step_ms": [round(float(x), 3) for x in step_ms],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "beam_search_kernel",
                         "kernel_ms": kern_avg_ms, "algorithmic_bytes_per_launch": int(alg_bytes)},
            "cpu_baseline": cpu_out,
            "knn": knn_out,
        }
        if sweep_table:
            line["config"]["sweep_csv"] = sweep_table
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
