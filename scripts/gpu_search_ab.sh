#!/bin/bash
# One gpurun call: A/B of the L2 fetch granularity for the search kernel, the per-phase trace of the host-buffer call,
# and the laion-like (512-d cosine) configuration.   usage: scripts/gpu_search_ab.sh [TAG]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
for g in default 32 64 128; do
  if [ "$g" = default ]; then unset RG_L2_FETCH; else export RG_L2_FETCH=$g; fi
  timeout 900 python bench.py --nb 10000000 --steps 10 --warmup 3 --L_pq 130 --no_cpu_baseline --no_knn \
      > $OUT/ab_l2fetch_${g}_$TAG.json 2> $OUT/ab_l2fetch_${g}_$TAG.err
  python - <<PY
import json
d=json.loads(open("$OUT/ab_l2fetch_${g}_$TAG.json").read().strip().splitlines()[-1])
print("l2fetch=$g", "qps", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],3))
PY
done
unset RG_L2_FETCH
RG_TRACE_E2E=1 timeout 600 python bench.py --nb 10000000 --steps 5 --warmup 3 --L_pq 130 --no_cpu_baseline --no_knn > $OUT/trace_e2e_$TAG.json 2> $OUT/trace_e2e_$TAG.err
grep "rg e2e" $OUT/trace_e2e_$TAG.err | tail -6
if [ -z "$NO_LAION" ]; then
  timeout 1500 python bench.py --nb 10000000 --dim 512 --dist cosine --steps 10 --warmup 3 --no_cpu_baseline \
      --sweep $OUT/sweep_laion_$TAG.csv > $OUT/bench_laion_$TAG.json 2> $OUT/bench_laion_$TAG.err; echo "laion rc=$?"
  tail -4 $OUT/bench_laion_$TAG.err; cat $OUT/bench_laion_$TAG.json
fi
