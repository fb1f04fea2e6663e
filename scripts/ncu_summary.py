#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` capture for profiles/: headline metrics per captured launch, top
warp-stall reasons, and the hottest CUDA source lines (stall samples / warp instructions aggregated per line from the
cuda,sass source page of the report itself — no cubin needed).
usage: ncu_summary.py <rep.ncu-rep> [top_lines] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum"]
print(f"# ncu --set full --clock-control none --import-source on: {rep.split('/')[-1]}")
for r in rows[2:]:
    print(f"\n## launch {r[0]}: {r[hdr.index('Kernel Name')][:150]}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:72s} {r[i]} {units[i]}")
    stalls = sorted(((float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr)
                     if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)[:6]
    print("  top warp-stall reasons (warps stalled per issue-active cycle):")
    for v, h in stalls:
        print(f"    {v:6.2f}  {h.split('stalled_')[1].split('_per_issue')[0]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = next((r for r in srows if r and r[0] == "Line No"), None)
if h is None or "# Samples" not in h:
    print("\n(no source correlation in this report)")
    sys.exit(0)
iS, iI = h.index("# Samples"), h.index("Instructions Executed")
agg, text, cur, fname, first_kernel_done = {}, {}, None, "?", False
for r in srows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        if first_kernel_done:
            break  # only the first captured launch
        first_kernel_done = True
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No" or len(r) <= iS:
        continue
    if r[0].strip():
        try:
            cur = (fname, int(r[0]))
            text[cur] = r[1]
        except ValueError:
            continue
    if len(r) > 2 and not r[2].strip():
        continue
    try:
        s, n = int(r[iS] or 0), int(r[iI] or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0])
    a[0] += n
    a[1] += s
ti, ts = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
print(f"\n## hottest CUDA source lines of the first captured launch ({ti} warp instructions, {ts} stall samples)")
for loc, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{s / ts * 100:5.1f}% smp {n / ti * 100:5.1f}% inst  {loc[0]}:{loc[1]}  {text.get(loc, '').strip()[:110]}")
