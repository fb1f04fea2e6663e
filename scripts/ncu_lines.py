#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --kernel-id :::N) with nvdisasm line info of the
same kernel, and print instruction / stall-sample shares per CUDA source line.
usage: ncu_lines.py <src.csv> <cubin> <mangled-kernel-substring> [top]"""
import csv, re, subprocess, sys

src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ii, sa = hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Address", "Kernel Name"):
        break  # next kernel instance
    if len(r) > sa:
        sass.append((r[1].strip(), int(r[ii] or 0), int(r[sa] or 0)))
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, active, inline = [], None, False, None
for ln in out.splitlines():
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        active = kname in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip())
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} disassembled vs {len(sass)} profiled instructions", file=sys.stderr)
agg = {}
for (txt, n, s), loc in zip(sass, lines):
    key = loc[:2] if loc else ("?", 0)
    a = agg.setdefault(key, [0, 0])
    a[0] += n
    a[1] += s
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"total warp instructions {ti}, stall samples {ts}")
srcs = {}
for (f, l), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in srcs:
        try:
            import glob
            p = glob.glob(f"/root/repo/**/{f}", recursive=True)[0]
            srcs[f] = open(p).read().splitlines()
        except Exception:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
    print(f"{s / ts * 100:5.1f}% smp {n / ti * 100:5.1f}% inst  {f}:{l}  {text}")
