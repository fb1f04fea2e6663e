#!/usr/bin/env python
"""SASS evidence for the shipped library: per kernel of roargraph_b200/librg_b200.so the resource usage (cuobjdump -res-usage)
and a histogram of the instruction mnemonics that prove the Blackwell paths (tcgen05 MMA = UTCHMMA / UTCHMMA.2CTA, TMA tiles =
UTMALDG, 1-D bulk copies = UBLKCP, TMEM loads = LDTM, mbarrier = SYNCS), plus the top mnemonics overall.
usage: sass_summary.py [kernel-name-substring ...] > profiles/rN_sass.txt   (needs only the CUDA toolkit, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "roargraph_b200", "librg_b200.so")
WANT = sys.argv[1:] or ["beam_search_kernel", "knn_tc_filter_pair_kernel", "knn_merge_fast_kernel", "topk_merge_packed_kernel",
                        "prune_kernel"]
PROOF = ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "SYNCS", "UTCATOMSWS", "FMNMX3", "REDUX", "LDGSTS",
         "HMMA", "STL", "LDL")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    print(f"# cuobjdump -sass / -res-usage of roargraph_b200/librg_b200.so ({os.path.getsize(LIB)} bytes), sm_100a")
    for k, cnt in kernels.items():
        dn = names.get(k, k)
        if not any(w in dn for w in WANT):
            continue
        total = sum(cnt.values())
        print(f"\n## {dn}\n   {usage.get(k, '')}\n   {total} SASS instructions")
        proof = {}
        for mn, c in cnt.items():
            for p in PROOF:
                if mn.startswith(p):
                    proof[mn] = c
        if proof:
            print("   Blackwell / notable: " + ", ".join(f"{m} x{c}" for m, c in sorted(proof.items())))
        print("   top: " + ", ".join(f"{m} x{c}" for m, c in cnt.most_common(14)))


if __name__ == "__main__":
    main()
