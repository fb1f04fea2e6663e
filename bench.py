#!/usr/bin/env python
"""bench.py — QPS of the batched RoarGraph beam search at recall@10 >= 0.9 on t2i-10M-shaped synthetic data, plus one
sub-record per further BASELINE.json configuration, all in ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--extras all|none|a,b,...]

Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): one "step" = one pass of SearchRoarGraph over the whole query batch
at the smallest L_pq of the reference's canonical sweep (README.md:118) whose recall@10 against exact ground truth is >= 0.9.
  N = 1  BASELINE configs[1]: 10 M x 200 IP, 10 000 queries top-10
  N > 1  BASELINE configs[4]: the same base sharded over the ranks (per-shard index, ONE packed all-gather + merge kernel per
         step), 100 000-query batch
  value     queries/s with queries and outputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same through the host-buffer entry (C-ABI rg_search_roargraph at N = 1; at N > 1 every rank uploads 1/N of the
            pinned host batch, NVLink all-gather of the rows, search, exchange, merge, D2H), wall clock around the call
  roofline  algorithmic gather bytes of the search kernel / its CUDA-event duration vs the measured HBM copy peak
  cpu_baseline  the reference's own SearchRoarGraph (compiled unmodified, oracle/_ref) on the host cores, same index
Sub-records (each carries its own workload, timing and roofline; a failure is recorded under its key, never fatal):
  batch100k / batch10k   the other batch size on the same index
  knn        BASELINE configs[3]: compute_groundtruth's exact_knn, 10 M base x 10 M training queries, K = 100, base sharded
             over the ranks (tcgen05 bf16 filter + exact fp32 re-rank, shared bound, packed exchange + merge per 151 552-query batch);
             at N = 1 with the reference's MKL path timed on a slice beside it (cpu_baseline)
  bipartite  SearchBipartiteGraph (A2) on a 1 M x 200 base with 3 M training queries (N = 1)
  laion      BASELINE configs[2]: 10 M x 512 cosine, 10 000 queries (N = 1)
  webvid     BASELINE configs[4]: 2.5 M x 512 cosine base-sharded, 100 000-query batch (N > 1)
  replicas   the comparison point of SURVEY.md 8(e): full index on every rank, the batch split over the ranks (N > 1)
  sharded_parity  N > 1: the unmodified reference run on every shard's index for 200 queries, merged on the host by
             (distance, id), against the GPU's merged ids
--impl reference prints the reference-CPU arm as the headline line instead (rank 0 only under torchrun).
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

ALL_EXTRAS = ("batch", "knn", "bipartite", "laion", "webvid", "replicas", "sharded_parity")


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
    ap.add_argument("--nq", type=int, default=0, help="query batch (default: 10 000 at N = 1, 100 000 at N > 1)")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--M_sq", type=int, default=100)
    ap.add_argument("--M_pjbp", type=int, default=35)
    ap.add_argument("--L_pjpq", type=int, default=500)
    ap.add_argument("--target_recall", type=float, default=0.9)
    ap.add_argument("--L_pq", type=int, default=0, help="fix L_pq instead of calibrating it on recall")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cache_dir", default=os.environ.get("RG_BENCH_CACHE", "/tmp/rg_bench_cache"))
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--extras", default="all", help="comma list of sub-records (%s), 'all' or 'none'" % ",".join(ALL_EXTRAS))
    ap.add_argument("--extras_div", type=int, default=1, help="divide the sizes of the sub-records' own workloads (shake-down runs)")
    ap.add_argument("--knn_nq", type=int, default=0, help="training queries of the kNN sub-record (default: nb, i.e. 10 M x 10 M)")
    ap.add_argument("--shard", default="base", choices=["base", "query"],
                    help="N > 1 headline: 'base' = base rows sharded, per-shard index, all-gather + merge (the contract); "
                         "'query' = full index replicated, the batch split over the ranks")
    ap.add_argument("--sweep", default="", help="also write the full L_pq sweep table (CSV) to this path (N = 1)")
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


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def load_peaks():
    pk = _peaks()
    if pk and pk.get("hbm_gbs"):
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak(burst=False):
    pk = _peaks()
    key = "bf16_tflops" if burst else "bf16_tflops_sustained"
    if pk and pk.get(key):
        return float(pk[key]), f"measured (MEASURED_PEAKS.json {key})"
    return (1650.0 if burst else 1400.0), "fallback (B200_PROFILING.md, " + ("burst)" if burst else "sustained)")


def workload_name(nb, dim, dist, nq, k, M_sq, M_pjbp, L_pjpq, ntrain):
    label = "t2i" if (dim == 200 and dist == "ip") else ("laion" if (dim == 512 and nb >= 5_000_000) else
                                                         "webvid" if dim == 512 else "synthetic")
    return (f"{label}-like synthetic {nb}x{dim} {dist}, {nq} queries top-{k}, RoarGraph M_sq={M_sq} "
            f"M_pjbp={M_pjbp} L_pjpq={L_pjpq} from {ntrain} training queries")


class Ctx:
    """What every leg needs: torch device, distributed layout, CLI arguments."""

    def __init__(self, a, dev, rank, local_rank, world):
        self.a, self.dev, self.rank, self.local_rank, self.world = a, dev, rank, local_rank, world
        self.flush = None

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def l2_flush(self, v: int):
        import torch
        if self.flush is None:
            self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.flush.fill_(v & 0xff)


def measure_search(cx: Ctx, w, d_q, gt_ids, k, steps, warmup, L_fix=0, target=0.9, mode="sharded", do_e2e=True,
                   sweep_path="", clocks=False):
    """Calibrate L_pq on recall, then time `steps` passes over the batch d_q (device-resident), the search kernel alone for
    the roofline, and the host-buffer path.  mode: "sharded" (ShardedRoarIndex over cx.world ranks; plain search at N = 1)
    or "replicas" (ReplicatedRoarIndex: the batch split over the ranks)."""
    import torch
    import torch.distributed as dist

    from roargraph_b200 import workload as W
    from roargraph_b200.sharded import ReplicatedRoarIndex, ShardedRoarIndex

    dev, world, rank = cx.dev, cx.world, cx.rank
    nq = d_q.shape[0]
    replicas = mode == "replicas" and world > 1
    sidx = ReplicatedRoarIndex(w.index) if replicas else ShardedRoarIndex(w.index, w.row_lo)
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_dists = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_cmps = torch.empty(nq, dtype=torch.int32, device=dev)
    d_hops = torch.empty(nq, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def run_dev(L):
        if replicas:
            return sidx.search_dev(d_q, k, L, d_ids, d_dists)
        return sidx.search_dev(d_q, k, L, d_ids, d_dists, d_cmps, d_hops)

    # ---- calibrate L_pq on recall (outside the timed region) ------------------------------------------------------
    L_star, rec_star, tried = L_fix, None, []
    if L_fix <= 0:
        for L in W.README_L_SWEEP:
            if L < k:
                continue
            run_dev(L)
            w.index.check(stream)
            rec = W.recall_at_k(d_ids.cpu().numpy().view(np.uint32), gt_ids, k)
            tried.append((L, rec))
            if rec >= target:
                L_star, rec_star = L, rec
                break
        if rec_star is None:
            L_star, rec_star = tried[-1]
            W.log(f"WARNING: recall {target} not reached inside the sweep; using L_pq={L_star} (recall {rec_star:.4f})")
    else:
        run_dev(L_star)
        w.index.check(stream)
        rec_star = W.recall_at_k(d_ids.cpu().numpy().view(np.uint32), gt_ids, k)
    L = L_star
    W.log(f"[{mode} x{world}] {nq} queries, L_pq={L}: recall@{k}={rec_star:.4f}")

    # ---- timed region: device-resident -----------------------------------------------------------------------------
    for _ in range(max(warmup, 3)):
        run_dev(L)
    w.index.check(stream)
    sampler = ClockSampler(cx.local_rank) if clocks else None
    if sampler is not None and sampler.ok:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    cx.barrier()
    torch.cuda.synchronize(dev)
    launches = 0
    for s in range(steps):
        cx.l2_flush(s)  # evict the previous step's rows from L2 (not timed)
        ev[s][0].record(stream)
        launches += run_dev(L)
        ev[s][1].record(stream)
    torch.cuda.synchronize(dev)
    cx.barrier()
    w.index.check(stream)
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    ms_per_step = cx.max_over_ranks(float(sum(step_ms))) / steps
    qps = nq / (ms_per_step / 1e3)

    # ---- the search kernel alone on this rank's queries (roofline) ---------------------------------------------------
    if replicas:
        per = -(-nq // world)
        d_q_local = d_q[min(rank * per, nq):min((rank + 1) * per, nq)]
    else:
        d_q_local = d_q
    nq_local = d_q_local.shape[0]
    loc_i = torch.empty((max(nq_local, 1), k), dtype=torch.int32, device=dev)
    loc_d = torch.empty((max(nq_local, 1), k), dtype=torch.float32, device=dev)
    kern_ms = []
    for s in range(min(steps, 10)):
        cx.l2_flush(s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if nq_local:
            w.index.search_dev(d_q_local, k, L, loc_i, loc_d, d_cmps[:nq_local], d_hops[:nq_local], stream=stream)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        kern_ms.append(e0.elapsed_time(e1))
    clk = None
    if sampler is not None:
        clk = sampler.stop() if sampler.ok else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                                                 "note": getattr(sampler, "err", "NVML unavailable")}
    # median: every 255th query of a slot clears its visited-tag array (VisitedList::reset), which lands in one launch
    kern_med = float(np.median(kern_ms))
    st = w.index.last_stats()   # of the kernel-only pass: this rank's queries on this rank's index
    ll = w.index.last_launch()
    alg_bytes = (st["cmps"] + nq_local) * w.d_base.shape[1] * 4 + st["adj_words"] * 4
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (kern_med / 1e3) / 1e9 if nq_local else 0.0
    shard_mb = w.d_base.shape[0] * w.d_base.shape[1] * 4 / 1e6
    l2_note = (f"; this rank's base shard is {shard_mb:.0f} MB and every row is gathered by many queries of the batch: part of the "
               f"algorithmic bytes is served by the 126 MB L2, so the fraction of the HBM peak can exceed what DRAM alone delivers"
               if shard_mb < 2000 else "")

    # ---- end to end through the host-buffer path ---------------------------------------------------------------------
    e2e = None
    e2e_ids = None
    if do_e2e:
        hq = d_q.cpu().pin_memory()
        e2e_ms = []
        n_e2e = warmup + steps
        if world == 1:
            h_ids = torch.empty((nq, k), dtype=torch.int32).pin_memory()
            h_d = torch.empty((nq, k), dtype=torch.float32).pin_memory()
            h_c = torch.empty(nq, dtype=torch.int32).pin_memory()
            h_h = torch.empty(nq, dtype=torch.int32).pin_memory()
            for s in range(n_e2e):
                cx.l2_flush(s)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                w.index.search_into(hq.numpy(), k, L, h_ids.numpy(), h_d.numpy(), h_c.numpy(), h_h.numpy())
                t1 = time.perf_counter()
                if s >= warmup:
                    e2e_ms.append((t1 - t0) * 1e3)
            h2d = hq.numel() * 4
            d2h = (h_ids.numel() + h_d.numel() + h_c.numel() + h_h.numel()) * 4
            e2e_ids = h_ids.numpy().view(np.uint32).copy()
            e2e_mean = float(np.mean(e2e_ms))
        else:
            h_ids = torch.empty((nq, k), dtype=torch.int32).pin_memory()
            h_d = torch.empty((nq, k), dtype=torch.float32).pin_memory()
            per = -(-nq // world)
            lo, hi = min(rank * per, nq), min((rank + 1) * per, nq)
            for s in range(n_e2e):
                cx.l2_flush(s)
                torch.cuda.synchronize(dev)
                dist.barrier()
                t0 = time.perf_counter()
                if replicas:  # this rank's slice H2D, search, all-gather, the whole result D2H
                    if hi > lo:
                        d_q[lo:hi].copy_(hq[lo:hi], non_blocking=True)
                    sidx.search_dev(d_q, k, L, d_ids, d_dists)
                    h_ids.copy_(d_ids, non_blocking=True)
                    h_d.copy_(d_dists, non_blocking=True)
                    w.index.check(stream)
                    e2e_ids = h_ids.numpy().view(np.uint32)
                else:
                    e2e_ids, _ = sidx.search(hq.numpy(), k, L)
                t1 = time.perf_counter()
                if s >= warmup:
                    e2e_ms.append((t1 - t0) * 1e3)
            e2e_mean = cx.max_over_ranks(float(np.mean(e2e_ms)))
            h2d = (hi - lo) * hq.shape[1] * 4  # per rank: 1/N of the batch over PCIe, the rest over NVLink
            d2h = nq * k * 8
        e2e = {"value": nq / (e2e_mean / 1e3), "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_mean, "step_ms": [round(float(x), 3) for x in e2e_ms],
               "recall": W.recall_at_k(e2e_ids, gt_ids, k)}

    sweep_table = None
    if sweep_path and rank == 0 and world == 1:
        rows = []
        for Ls in W.README_L_SWEEP:
            if Ls < k:
                continue
            ts = []
            for rep in range(3):
                cx.l2_flush(rep)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                w.index.search_dev(d_q, k, Ls, d_ids, d_dists, d_cmps, d_hops, stream=stream)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1))
            s2 = w.index.last_stats()
            rec = W.recall_at_k(d_ids.cpu().numpy().view(np.uint32), gt_ids, k)
            ms = min(ts)
            gbs = ((s2["cmps"] + nq) * w.d_base.shape[1] * 4 + s2["adj_words"] * 4) / (ms / 1e3) / 1e9
            rows.append((Ls, nq / (ms / 1e3), s2["cmps"] / nq, ms / nq, rec, s2["hops"] / nq, gbs, gbs / peak))
        os.makedirs(os.path.dirname(os.path.abspath(sweep_path)), exist_ok=True)
        with open(sweep_path, "w") as f:
            f.write("L_pq,qps,avg_cmps,ms_per_query,recall,avg_hops,gather_GBps,hbm_frac\n")  # CLI columns + roofline
            for r in rows:
                f.write(",".join(f"{x:.6g}" for x in r) + "\n")
        sweep_table = sweep_path

    return {"L_pq": L, "recall": rec_star, "value": qps, "ms_per_step": ms_per_step, "steps": steps,
            "step_ms": [round(float(x), 3) for x in step_ms], "launches": int(launches), "nq": nq,
            "avg_cmps": st["cmps"] / max(nq_local, 1), "avg_hops": st["hops"] / max(nq_local, 1), "launch": ll,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "kernel": "beam_search_kernel", "kernel_ms": kern_med,
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "note": "this rank's queries on this rank's index, search kernel alone" + l2_note},
            "e2e": e2e, "e2e_ids": e2e_ids, "clocks": clk, "sweep_csv": sweep_table, "d_ids": d_ids}


def time_reference_cpu(cx, w, q_host, gt_ids, metric, k, L, steps, warmup, budget_s):
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
    os.makedirs(cx.a.cache_dir, exist_ok=True)
    ipath = os.path.join(cx.a.cache_dir, f"ref_index_{os.getpid()}.rg")
    F.write_projection_index(ipath, ep, row_off, nbrs)
    del row_off, nbrs
    base_host = w.d_base.cpu().numpy()
    h = R.index_from_memory(base_host, ipath, "l2" if metric == "l2" else "ip", cores)
    os.remove(ipath)
    log(f"reference index in host memory ({time.time() - t0:.1f}s), {cores} host threads")
    q = q_host
    # size the sample on a small slice (this call also does the CLI's 100-query serial warm-up)
    r0 = R.search_roargraph(h, q[:min(len(q), 1000)], k, L, cores, warmup=True)
    qps_est = min(len(q), 1000) / max(r0["wall_s"], 1e-6)
    n_sample = int(min(len(q), max(1000, qps_est * budget_s / max(1, steps + warmup))))
    for _ in range(warmup):
        R.search_roargraph(h, q[:n_sample], k, L, cores, warmup=False)
    walls, res = [], None
    for _ in range(steps):
        res = R.search_roargraph(h, q[:n_sample], k, L, cores, warmup=False)
        walls.append(res["wall_s"])
    R.index_free(h)
    total = float(sum(walls))
    rec = recall_at_k(res["ids"], gt_ids[:n_sample], k)
    return {"value": n_sample * steps / total, "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": f"first {n_sample} of {len(q)} queries at L_pq={L} per step, {steps} timed passes after {warmup} "
                      f"untimed ones and the CLI's 100-query serial warm-up; recall@{k}={rec:.4f}",
            "ms_per_step": total / steps * 1e3, "step_ms": [round(x * 1e3, 3) for x in walls], "n_sample": n_sample,
            "res": res}


# ---- sub-records ---------------------------------------------------------------------------------------------------------

def leg_batch(cx, w, k, L_main, nq_other):
    """The other batch size of the pair {10 000 (configs[1]), 100 000 (configs[4])} on the same index."""
    d_q = w.make_queries(nq_other)
    gt = w.exact_topk(d_q, k)
    m = measure_search(cx, w, d_q, gt, k, steps=min(cx.a.steps, 5), warmup=3, target=cx.a.target_recall)
    return {"workload": f"same index, {nq_other}-query batch", "L_pq": m["L_pq"], "recall": m["recall"], "value": m["value"],
            "unit": "queries/s", "ms_per_step": m["ms_per_step"], "steps": m["steps"], "roofline": m["roofline"],
            "e2e": {kk: vv for kk, vv in m["e2e"].items() if kk != "step_ms"}, "avg_cmps": m["avg_cmps"]}


def leg_knn(cx, w):
    """BASELINE configs[3]: exact_knn of nq training queries against the (sharded) base, K = 100; the whole job, inputs
    resident: base preparation (norms, bf16 copy) + every 151 552-query batch (seed pass, filter, re-rank/merge, packed
    all-gather + merge of the own slice).  Wall clock between device synchronisations, max over ranks."""
    import torch

    import roargraph_b200 as rg
    from roargraph_b200.sharded import ShardedKnn
    from roargraph_b200.workload import log

    a, dev, world = cx.a, cx.dev, cx.world
    K = 100
    nq = a.knn_nq or int(w.nd_total) // a.extras_div
    metric = "l2" if a.dist == "l2" else "ip"
    free_b, _ = torch.cuda.mem_get_info(dev)
    need = nq * (w.d_base.shape[1] * 4 + K * 8) + (8 << 30)
    if need > free_b:  # never drive a box out of memory: shrink the job and say so
        nq = int((free_b - (8 << 30)) // (w.d_base.shape[1] * 4 + K * 8)) // 151552 * 151552
    t0 = time.time()
    q = w.make_queries(nq)
    out_i = torch.empty((nq, K), dtype=torch.int32, device=dev)
    out_d = torch.empty((nq, K), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    log(f"kNN leg: {nq} training queries on device ({time.time() - t0:.1f}s)")
    # warm-up: module load, kernel attributes, workspace growth, NCCL channels
    sk = ShardedKnn(w.d_base, metric, w.row_lo)
    sk.search(q[:min(nq, 2 * 151552)], K, slice_only=True, out=(out_i[:min(nq, 2 * 151552)], out_d[:min(nq, 2 * 151552)]))
    sk.close()
    torch.cuda.synchronize(dev)
    cx.barrier()
    t0 = time.perf_counter()
    sk = ShardedKnn(w.d_base, metric, w.row_lo)          # base preparation is part of the job
    sk.search(q, K, slice_only=True, out=(out_i, out_d))
    torch.cuda.synchronize(dev)
    dt = cx.max_over_ranks(time.perf_counter() - t0)
    # per-phase device times of one batch (a separate, synchronising pass)
    rg.knn_trace(1)
    rg.lib().rg_knn_profile_reset()
    nb1 = min(nq, 151552)
    sk.search(q[:nb1], K, slice_only=True, out=(out_i[:nb1], out_d[:nb1]))
    torch.cuda.synchronize(dev)
    phases = rg.knn_profile(reset=True)
    rg.knn_trace(-1)
    # spot check against the fp32 SIMT filter (no bf16 anywhere) on a sample, through the same sharded exchange
    ns = min(nq, 2048)
    os.environ["RG_KNN_MODE"] = "simt"
    try:
        sk2 = ShardedKnn(w.d_base, metric, w.row_lo)
        si, sd = sk2.search(q[:ns], K)
        sk2.close()
    finally:
        os.environ.pop("RG_KNN_MODE", None)
    ti, td = sk.search(q[:ns], K)
    torch.cuda.synchronize(dev)
    same = float((si == ti).float().mean().item())
    max_rel = float(((sd - td).abs() / sd.abs().clamp_min(1e-6)).max().item())
    sk.close()
    del q, out_i, out_d
    rg.knn_release_workspace()
    torch.cuda.empty_cache()
    pairs = float(nq) * float(w.nd_total)
    peak, src = load_tensor_peak()
    burst, _ = load_tensor_peak(burst=True)
    tf = 2.0 * pairs * a.dim / dt / 1e12
    return {"metric": "brute-force kNN pairs/s", "value": pairs / dt, "unit": "pairs/s",
            "workload": f"compute_groundtruth exact_knn: {int(w.nd_total)} base x {nq} training queries, d={a.dim}, K={K}, {metric}, "
                        f"base rows sharded over {world} rank(s)",
            "nq": nq, "nb": int(w.nd_total), "dim": a.dim, "K": K, "dist": metric, "seconds": dt, "tflops": tf,
            "n_gpus": world, "scaling": "strong",
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak * world, "unit": "TFLOP/s", "frac": tf / (peak * world),
                         "peak_source": src, "peak_burst": burst * world, "frac_of_burst": tf / (burst * world),
                         "kernel": "knn_tc_filter_pair_kernel (tcgen05 bf16 filter) + knn_merge_fast_kernel (fp32 re-rank)",
                         "flops": "2 * nq * nb * d with the true d (the MMA runs K = 208)"},
            "phases_ms_one_batch": {**phases, "batch": nb1, "note": "device time per phase of one 151 552-query batch on this "
                                    "rank (separate synchronising pass); the exchange (packed all-to-all of the slices + merge) runs on a "
                                    "side stream under the next batch's filter"},
            "check": {"sample": ns, "ids_equal_to_fp32_simt_path": same, "max_rel_distance_diff": max_rel,
                      "note": "independent fp32 path with a different accumulation order: distances agree to ~1e-6 relative and "
                              "the ids differ only where two rows are closer than that (near-ties swap places); the reference "
                              "itself is compared in cpu_baseline.ids_equal_to_gpu_fraction and tests/test_gpu_knn.py (MKL, 200 k x 200 k)"},
            "note": "wall clock of the whole job between device synchronisations (max over ranks): base preparation, every "
                    "batch's conversion, tau seed pass, filter, re-rank/merge, unpack, packed all-to-all + merge of the rank's "
                    "slice; queries and results resident in HBM"}


def leg_knn_cpu_baseline(cx, w, budget_pairs=2.5e10):
    """The reference's compute_groundtruth (MKL SGEMM + heap scan, compiled unmodified: oracle/_ref/ref_compute_groundtruth) on
    a slice of the same workload, on the box's host cores; pairs/s of the slice, and the GPU's ids on that slice against it."""
    import tempfile

    import torch

    import roargraph_b200 as rg
    from oracle.oracle import HERE as ORACLE_DIR, Oracle, run_ref_compute_groundtruth
    from roargraph_b200 import formats as F

    exe = os.path.join(ORACLE_DIR, "_ref", "ref_compute_groundtruth")
    if not os.path.exists(exe):
        return {"value": None, "kind": "reference", "sample": "oracle/_ref/ref_compute_groundtruth not available"}
    a = cx.a
    nb_s = min(int(w.d_base.shape[0]), 1_000_000)
    nq_s = max(1024, int(budget_pairs // nb_s) // 1024 * 1024)
    K = 100
    metric = "l2" if a.dist == "l2" else "ip"
    base = w.d_base[:nb_s, :a.dim].contiguous()
    q = w.d_train_sample[:nq_s, :a.dim].contiguous()
    nq_s = q.shape[0]
    d = tempfile.mkdtemp(prefix="rg_knn_ref_", dir=a.cache_dir if os.path.isdir(a.cache_dir) else None)
    bf, qf, gf = (os.path.join(d, n) for n in ("base.fbin", "q.fbin", "gt.bin"))
    F.write_fbin(bf, base.cpu().numpy())
    F.write_fbin(qf, q.cpu().numpy())
    t0 = time.perf_counter()
    run_ref_compute_groundtruth(bf, qf, gf, K, "mips" if metric == "ip" else "l2")
    dt = time.perf_counter() - t0
    ref_i, ref_d = F.read_truthset(gf)
    for f in (bf, qf, gf):
        os.remove(f)
    os.rmdir(d)
    gi = torch.empty((nq_s, K), dtype=torch.int32, device=cx.dev)
    gd = torch.empty((nq_s, K), dtype=torch.float32, device=cx.dev)
    rg.knn_dev(base, q, K, metric, gi, gd)
    torch.cuda.synchronize(cx.dev)
    same = float((gi.cpu().numpy().view(np.uint32) == ref_i).mean())
    cores = Oracle().num_procs()
    return {"value": nq_s * float(nb_s) / dt, "unit": "pairs/s", "cores": cores, "kind": "reference",
            "sample": f"{nb_s} base rows x {nq_s} training queries, K={K}, {metric}: wall clock of the whole tool run (file load, "
                      f"MKL SGEMM batches of 1 024 queries, heap scan, file write)", "seconds": dt,
            "ids_equal_to_gpu_fraction": same,
            "note": "MKL = oneMKL 2024.2 inside torch's libtorch_cpu.so (the reference's MKL is un-vendored and unpinned)"}


def leg_bipartite(cx):
    """A2 SearchBipartiteGraph: raw bipartite graph of qbaseNNbipartite (src/index_bipartite.cpp:235-280, M_pjbp = 35) over a
    1 M x 200 base with 3 M training queries (three per base row, so that the weak baseline reaches recall 0.9 inside the
    sweep: with the README's 10 % only 4 % of the base nodes own an out-edge and recall saturates at 0.3), 10 000 queries,
    10 uniformly random entry points per query (the reference draws them from an unseeded random_device, :287-294)."""
    import tempfile

    import torch

    import roargraph_b200 as rg
    from roargraph_b200 import formats as F, synth, workload as W

    a, dev = cx.a, cx.dev
    nb, ntrain, nq, k, M = 1_000_000 // a.extras_div, 3_000_000 // a.extras_div, 10_000, a.k, a.M_pjbp
    base, train, test = synth.make_torch(nb, ntrain, nq, a.dim, dev, seed=a.seed + 1)
    gt_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    gt_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    rg.knn_dev(base, test, k, "ip", gt_i, gt_d)
    kn_i = torch.empty((ntrain, M), dtype=torch.int32, device=dev)
    kn_d = torch.empty((ntrain, M), dtype=torch.float32, device=dev)
    rg.knn_dev(base, train, M, "ip", kn_i, kn_d)
    torch.cuda.synchronize(dev)
    gt = gt_i.cpu().numpy().view(np.uint32)
    row_off, nbrs = W.bipartite_csr(kn_i.cpu().numpy().view(np.uint32), nb, M)
    covered = float((np.diff(row_off[:nb + 1].astype(np.int64)) > 0).mean())
    del kn_i, kn_d, train
    rg.knn_release_workspace()
    idx = rg.RoarIndex(base, "ip")
    idx.set_bipartite(row_off, nbrs)
    # 10 entry points per query, uniform over the base rows like the reference's (src/index_bipartite.cpp:287-294)
    init = np.random.default_rng(7).integers(0, nb, (nq, 10)).astype(np.uint32)
    d_init = torch.from_numpy(init.view(np.int32)).to(dev)
    d_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_c = torch.empty(nq, dtype=torch.int32, device=dev)
    d_h = torch.empty(nq, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    L_star, rec_star = None, None
    for L in W.README_L_SWEEP:
        if L < k:
            continue
        idx.search_bipartite_dev(test, k, L, d_i, d_d, d_c, d_h, d_init_ids=d_init, stream=stream)
        idx.check(stream)
        rec = W.recall_at_k(d_i.cpu().numpy().view(np.uint32), gt, k)
        L_star, rec_star = L, rec
        if rec >= a.target_recall:
            break
    L = L_star
    ts = []
    for s in range(3 + 5):
        cx.l2_flush(s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        idx.search_bipartite_dev(test, k, L, d_i, d_d, d_c, d_h, d_init_ids=d_init, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if s >= 3:
            ts.append(e0.elapsed_time(e1))
    st = idx.last_stats()
    ms = float(np.median(ts))
    alg = (st["cmps"] + nq * 10) * base.shape[1] * 4 + st["adj_words"] * 4
    peak, src = load_peaks()
    out = {"path": "SearchBipartiteGraph (A2)", "workload": f"{nb}x{a.dim} ip, {ntrain} training queries, M_pjbp={M}, {nq} queries "
           f"top-{k}, 10 entry points per query; {covered:.3f} of the base nodes own an out-edge",
           "L_pq": L, "recall": rec_star, "value": nq / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms, "steps": 5,
           "avg_cmps": st["cmps"] / nq, "avg_hops": st["hops"] / nq,
           "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (ms / 1e3) / 1e9 / peak, "peak_source": src, "kernel": "beam_search_kernel (bipartite mode)",
                        "kernel_ms": ms, "algorithmic_bytes_per_launch": int(alg)}}
    # the reference's own SearchBipartiteGraph on the host cores over the same graph: timing (stock library, unseeded entry
    # points) and id parity (the library compiled with a fixed random_device word per query, entry points handed to the GPU)
    try:
        from oracle.oracle import Reference, have_detrand_reference, have_reference
        if have_reference() and not a.no_cpu_baseline:
            d = tempfile.mkdtemp(prefix="rg_bip_ref_", dir=a.cache_dir if os.path.isdir(a.cache_dir) else None)
            bf, gf = os.path.join(d, "base.fbin"), os.path.join(d, "bp.rg")
            F.write_fbin(bf, base.cpu().numpy())
            F.write_bipartite_index(gf, row_off, nbrs)
            qh = test.cpu().numpy()
            R = Reference()
            cores = R.num_procs()
            h = R.bipartite_load(bf, gf, "ip", base.shape[1], ntrain, cores)
            r0 = R.search_bipartite(h, qh[:256], k, L, cores)
            n_s = int(min(nq, max(256, 256 / max(r0["wall_s"], 1e-6) * 8.0)))
            r = R.search_bipartite(h, qh[:n_s], k, L, cores)
            R.index_free(h)
            out["cpu_baseline"] = {"value": n_s / r["wall_s"], "unit": "queries/s", "cores": cores, "kind": "reference",
                                   "sample": f"first {n_s} queries at L_pq={L}, one pass, the reference's own unseeded entry points"}
            if have_detrand_reference():
                Rd = Reference(detrand=True)
                h = Rd.bipartite_load(bf, gf, "ip", base.shape[1], ntrain, cores)
                want = Rd.search_bipartite(h, qh[:500], k, L, cores, seed_base=99)
                Rd.index_free(h)
                got = idx.search_bipartite(qh[:500], k, L, init_ids=want["init_ids"])
                out["ids_identical_to_reference"] = bool((got[0] == want["ids"]).all() and (got[2] == want["cmps"]).all()
                                                         and (got[1].view(np.uint32) == want["dists"].view(np.uint32)).all())
            os.remove(bf)
            os.remove(gf)
            os.rmdir(d)
    except Exception as e:  # noqa: BLE001
        out["cpu_baseline"] = {"value": None, "kind": "reference", "sample": f"failed: {e!r}"}
    idx.close()
    return out


def leg_other_workload(cx, name, nb, dim, dist, nq, ref_check):
    """A second search workload through the same measurement (laion: configs[2]; webvid: configs[4])."""
    import torch

    import roargraph_b200 as rg
    from roargraph_b200 import workload as W

    a = cx.a
    ntrain = max(1000, nb // 10)
    n_shards = cx.world
    w = W.build_search_workload(nb, dim, dist, ntrain, nq, a.k, a.M_sq, a.M_pjbp, a.L_pjpq, a.seed, cx.local_rank,
                                shard=cx.rank if n_shards > 1 else 0, n_shards=n_shards, cache_dir=a.cache_dir)
    m = measure_search(cx, w, w.d_queries, w.gt_ids, a.k, steps=min(a.steps, 10), warmup=3, target=a.target_recall)
    out = {"workload": workload_name(nb, dim, dist, nq, a.k, a.M_sq, a.M_pjbp, a.L_pjpq, ntrain)
           + (f", base rows over {n_shards} ranks" if n_shards > 1 else ""),
           "L_pq": m["L_pq"], "recall": m["recall"], "value": m["value"], "unit": "queries/s", "ms_per_step": m["ms_per_step"],
           "steps": m["steps"], "roofline": m["roofline"], "e2e": {kk: vv for kk, vv in m["e2e"].items() if kk != "step_ms"},
           "avg_cmps": m["avg_cmps"], "avg_hops": m["avg_hops"], "launch": m["launch"], "setup": w.info, "n_gpus": cx.world}
    if ref_check and cx.rank == 0 and cx.world == 1 and not a.no_cpu_baseline:
        try:
            cpu = time_reference_cpu(cx, w, w.queries_host, w.gt_ids, dist, a.k, m["L_pq"], steps=2, warmup=0, budget_s=8.0)
            n_s = cpu["n_sample"]
            out["cpu_baseline"] = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            out["cpu_baseline"]["ids_identical_to_gpu"] = bool((cpu["res"]["ids"] == m["e2e_ids"][:n_s]).all())
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "kind": "reference", "sample": f"failed: {e!r}"}
    w.index.close()
    del w, m
    rg.knn_release_workspace()
    torch.cuda.empty_cache()
    return out


def leg_replicas(cx, nq, k, L_hint):
    """Query-sharded replicas: the FULL index on every rank (same seed -> same data; the deterministic device build gives
    every rank the same graph), the batch split over the ranks, one packed all-gather, no merge."""
    import torch

    import roargraph_b200 as rg
    from roargraph_b200 import workload as W

    a = cx.a
    w = W.build_search_workload(a.nb, a.dim, a.dist, a.ntrain, nq, k, a.M_sq, a.M_pjbp, a.L_pjpq, a.seed, cx.local_rank,
                                shard=0, n_shards=1, cache_dir=a.cache_dir)
    m = measure_search(cx, w, w.d_queries, w.gt_ids, k, steps=min(a.steps, 5), warmup=3, target=a.target_recall, mode="replicas")
    out = {"workload": workload_name(a.nb, a.dim, a.dist, nq, k, a.M_sq, a.M_pjbp, a.L_pjpq, a.ntrain)
           + f", full index on each of {cx.world} ranks, the batch split over the ranks, one all-gather",
           "L_pq": m["L_pq"], "recall": m["recall"], "value": m["value"], "unit": "queries/s", "ms_per_step": m["ms_per_step"],
           "steps": m["steps"], "roofline": m["roofline"], "e2e": {kk: vv for kk, vv in m["e2e"].items() if kk != "step_ms"},
           "n_gpus": cx.world, "scaling": "strong"}
    w.index.close()
    del w, m
    rg.knn_release_workspace()
    torch.cuda.empty_cache()
    return out


def leg_sharded_parity(cx, w, d_q, k, L, gpu_ids, n=200):
    """N > 1: every rank runs the UNMODIFIED reference SearchRoarGraph on its own shard's index for the first n queries
    (cores / N host threads each); the per-shard lists (global ids) are all-gathered and merged on the host by (distance, id)
    — the oracle SURVEY.md 8(e) names for the sharded path — and compared with the GPU's merged ids."""
    import torch
    import torch.distributed as dist

    from oracle.oracle import Reference
    from roargraph_b200 import formats as F

    R = Reference()
    threads = max(1, R.num_procs() // cx.world)
    ep, row_off, nbrs = w.index.get_graph()
    os.makedirs(cx.a.cache_dir, exist_ok=True)
    ipath = os.path.join(cx.a.cache_dir, f"ref_shard_{cx.rank}_{os.getpid()}.rg")
    F.write_projection_index(ipath, ep, row_off, nbrs)
    base_host = w.d_base.cpu().numpy()
    h = R.index_from_memory(base_host, ipath, "l2" if cx.a.dist == "l2" else "ip", threads)
    os.remove(ipath)
    q = d_q[:n].cpu().numpy()
    r = R.search_roargraph(h, q, k, L, threads)
    R.index_free(h)
    loc_i = torch.from_numpy((r["ids"].astype(np.int64) + w.row_lo).astype(np.int32)).to(cx.dev)
    loc_d = torch.from_numpy(r["dists"]).to(cx.dev)
    all_i = torch.empty((cx.world * n, k), dtype=torch.int32, device=cx.dev)
    all_d = torch.empty((cx.world * n, k), dtype=torch.float32, device=cx.dev)
    dist.all_gather_into_tensor(all_i, loc_i)
    dist.all_gather_into_tensor(all_d, loc_d)
    ai = all_i.cpu().numpy().view(np.uint32).reshape(cx.world, n, k).transpose(1, 0, 2).reshape(n, cx.world * k)
    ad = all_d.cpu().numpy().reshape(cx.world, n, k).transpose(1, 0, 2).reshape(n, cx.world * k)
    order = np.lexsort((ai, ad), axis=1)[:, :k]
    want = np.take_along_axis(ai, order, 1)
    got = gpu_ids[:n]
    return {"queries": n, "L_pq": L, "ids_identical_to_reference_merge": bool((want == got).all()),
            "mismatching_queries": int((want != got).any(axis=1).sum()),
            "oracle": "reference SearchRoarGraph (oracle/_ref, unmodified) on each shard's index + host merge by (distance, id)"}


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
    eff_world = world if a.impl == "b200" else 1
    cx = Ctx(a, dev, rank, local_rank, eff_world)
    k = a.k
    # headline batch: configs[1] (10 000 queries) on one GPU, configs[4] (100 000-query batch) on several
    nq = a.nq or (10_000 if max(a.gpus, world) == 1 else 100_000)
    replicas_headline = eff_world > 1 and a.shard == "query"
    if replicas_headline:
        n_shards = 1
    w = W.build_search_workload(a.nb, a.dim, a.dist, a.ntrain, nq, k, a.M_sq, a.M_pjbp, a.L_pjpq, a.seed, local_rank,
                                shard=rank if n_shards > 1 else 0, n_shards=n_shards, cache_dir=a.cache_dir)
    wname = workload_name(a.nb, a.dim, a.dist, nq, k, a.M_sq, a.M_pjbp, a.L_pjpq, a.ntrain)
    config = {"workload": wname, "k": k, "target_recall": a.target_recall}

    if a.impl == "reference":
        # operating point: the L_pq the GPU search needs for the recall target on this (whole, unsharded) index
        m = measure_search(cx, w, w.d_queries, w.gt_ids, k, steps=1, warmup=1, L_fix=a.L_pq, target=a.target_recall, do_e2e=False)
        cpu = time_reference_cpu(cx, w, w.queries_host, w.gt_ids, a.dist, k, m["L_pq"], steps=a.steps, warmup=a.warmup, budget_s=90.0)
        line = {
            "impl": "reference", "metric": f"QPS@recall{k}>={a.target_recall}", "value": cpu["value"], "unit": "queries/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": cpu["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "operating_point": {"L_pq": m["L_pq"], "recall": m["recall"],
                                "note": "unmodified reference SearchRoarGraph on the host cores over the WHOLE index; inputs (base, "
                                        "index, ground truth) were produced on the GPU before the timed region"},
            "cpu_baseline": {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cpu["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "step_ms": cpu["step_ms"],
        }
        print(json.dumps(line), flush=True)
        return 0

    m = measure_search(cx, w, w.d_queries, w.gt_ids, k, steps=a.steps, warmup=a.warmup, L_fix=a.L_pq, target=a.target_recall,
                       mode="replicas" if replicas_headline else "sharded", sweep_path=a.sweep, clocks=True)
    L = m["L_pq"]
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of this very workload, if there is one
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{a.nb}x{a.dim}:{a.dist}:L{L}", {}).get("traffic_bytes") if eff_world == 1 else None
    except OSError:
        pass
    roofline = dict(m["roofline"])
    roofline["traffic"] = traffic
    cpu_out = None
    if rank == 0 and eff_world == 1 and not a.no_cpu_baseline:
        try:
            cpu = time_reference_cpu(cx, w, w.queries_host, w.gt_ids, a.dist, k, L, steps=5, warmup=1, budget_s=20.0)
            n_s = cpu["n_sample"]
            cpu_out = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            cpu_out["ids_identical_to_gpu"] = bool((cpu["res"]["ids"] == m["e2e_ids"][:n_s]).all())
        except Exception as e:  # noqa: BLE001
            cpu_out = {"value": None, "unit": "queries/s", "cores": None, "kind": "reference", "sample": f"failed: {e!r}"}

    line = {
        "metric": f"QPS@recall{k}>={a.target_recall}", "value": m["value"], "unit": "queries/s", "n_gpus": eff_world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "operating_point": {"L_pq": L, "recall": m["recall"], "e2e_recall": m["e2e"]["recall"],
                            "sharding": (f"full index on each of {eff_world} ranks, the batch split over the ranks, one all-gather"
                                         if replicas_headline else
                                         f"base rows over {eff_world} ranks, per-shard index searched at the per-shard L_pq that "
                                         f"reaches the recall target after the merge, ONE packed all-gather + merge kernel per step"
                                         if eff_world > 1 else "none"),
                            "l2": "256 MiB fill between steps (outside the timed events); touched rows >> 126 MB L2",
                            "avg_cmps": m["avg_cmps"], "avg_hops": m["avg_hops"], "launch": m["launch"], "setup": w.info},
        "clocks": m["clocks"],
        "e2e": {kk: vv for kk, vv in m["e2e"].items() if kk != "recall"},
        "gpu_launches": m["launches"],
        "step_ms": m["step_ms"],
        "roofline": roofline,
        "cpu_baseline": cpu_out,
    }
    if m["sweep_csv"]:
        line["operating_point"]["sweep_csv"] = m["sweep_csv"]

    # ---- sub-records: never fatal, the headline above is already measured ------------------------------------------------
    want = set(ALL_EXTRAS) if a.extras == "all" else set() if a.extras == "none" else set(a.extras.split(","))

    def run_leg(key, fn, *args, **kw):
        t0 = time.time()
        try:
            out = fn(*args, **kw)
        except Exception as e:  # noqa: BLE001
            import traceback
            W.log(f"sub-record {key} failed: {e!r}\n{traceback.format_exc()}")
            out = {"error": repr(e)}
        if isinstance(out, dict):
            out["leg_seconds"] = round(time.time() - t0, 1)
        W.log(f"sub-record {key}: {time.time() - t0:.1f}s")
        return out

    try:
        if "sharded_parity" in want and eff_world > 1 and not replicas_headline:
            gpu_ids = m["d_ids"].cpu().numpy().view(np.uint32)
            line["sharded_parity"] = run_leg("sharded_parity", leg_sharded_parity, cx, w, w.d_queries, k, L, gpu_ids)
        if "batch" in want and not replicas_headline:
            other = (100_000 if nq <= 10_000 else 10_000) // (a.extras_div if a.extras_div > 1 and nq <= 10_000 else 1)
            line["batch100k" if other == 100_000 else "batch10k"] = run_leg("batch", leg_batch, cx, w, k, L, other)
        if "knn" in want and not replicas_headline:
            line["knn"] = run_leg("knn", leg_knn, cx, w)
            if rank == 0 and eff_world == 1 and not a.no_cpu_baseline and isinstance(line["knn"], dict):
                line["knn"]["cpu_baseline"] = run_leg("knn_cpu_baseline", leg_knn_cpu_baseline, cx, w)
        # the headline workload is no longer needed: give its HBM back before the other workloads are built
        w.index.close()
        del w, m
        rg.knn_release_workspace()
        torch.cuda.empty_cache()
        if "bipartite" in want and eff_world == 1:
            line["bipartite"] = run_leg("bipartite", leg_bipartite, cx)
        if "webvid" in want and eff_world > 1:
            line["webvid"] = run_leg("webvid", leg_other_workload, cx, "webvid", 2_500_000 // a.extras_div, 512, "cosine", 100_000 // a.extras_div, False)
        if "replicas" in want and eff_world > 1 and not replicas_headline:
            line["replicas"] = run_leg("replicas", leg_replicas, cx, nq, k, L)
        if "laion" in want and eff_world == 1:
            line["laion"] = run_leg("laion", leg_other_workload, cx, "laion", 10_000_000 // a.extras_div, 512, "cosine", 10_000, True)
    finally:
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
