#!/usr/bin/env python
"""Text summary of an `ncu --set full` capture for profiles/: headline metrics per captured launch + per-line hot spots.
usage: summarize_ncu.py <rep.ncu-rep> <cubin> <mangled-kernel-substring> > profiles/<name>.txt"""
import csv, io, subprocess, sys

rep, cubin, kname = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# ncu --set full --clock-control none: {rep}")
for r in rows[2:]:
    print(f"\n## launch {r[0]}: {r[hdr.index('Kernel Name')]}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {r[i]} {units[i]}")
    stalls = sorted(((float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr)
                     if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)[:6]
    print("  top warp-stall reasons (warps stalled per issue-active cycle):")
    for v, h in stalls:
        print(f"    {v:6.2f}  {h.split('stalled_')[1].split('_per_issue')[0]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
print("\n## hot CUDA source lines of the first captured launch (stall samples / warp instructions)")
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("summarize_ncu.py", "ncu_lines.py"), "/tmp/_src.csv", cubin, kname, "25"])
