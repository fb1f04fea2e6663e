#!/bin/bash
# search parity tests, the per-call phase trace of the host-buffer entry, then the L_pq sweep
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_build.py -x -q -m gpu > $OUT/pytest_search_$TAG.log 2>&1; tail -n 2 $OUT/pytest_search_$TAG.log
RG_TRACE_E2E=1 timeout 900 python bench.py --nb 10000000 --steps 40 --warmup 3 --L_pq 130 --no_cpu_baseline --no_knn \
    --sweep $OUT/sweep_$TAG.csv > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
grep "rg e2e" $OUT/bench_$TAG.err | awk '{print $7}' | sort -n | awk '{a[NR]=$1} END {print "e2e search ms: min", a[1], "median", a[int((NR+1)/2)], "p90", a[int(NR*0.9)], "max", a[NR], "n", NR}'
python -c "
import json
d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e'])
"
cut -d, -f1,2,8 $OUT/sweep_$TAG.csv | tr '\n' ' '; echo
