#!/bin/bash
# One gpurun call: GPU test-suite, a bench run at $NB base rows, the ncu launch list and one full capture of the search kernel.
# usage: scripts/gpu_check.sh [NB] [TAG] [TESTS]      (run from the repo root on the GPU box)
NB=${1:-1000000}
TAG=${2:-r1}
TESTS=${3:-tests}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/gpu_$TAG.txt
if [ "$TESTS" != "none" ]; then
  timeout 1500 python -m pytest $TESTS -q -m gpu > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
  tail -5 $OUT/pytest_$TAG.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
timeout 1200 python bench.py --nb $NB --steps 20 --warmup 3 --sweep $OUT/sweep_$TAG.csv > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -3 $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json
L=$(python -c "import json;print(json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])['config']['L_pq'])" 2>/dev/null || echo 50)
# launch list of the same command (graph comes from the cache the first run left in /tmp): kernel SHARES, not absolutes
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'beam_search|topk_merge' -c 40 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --nb $NB --steps 5 --warmup 3 --L_pq $L --no_cpu_baseline > $OUT/ncu_bench_$TAG.log 2>&1
# one full capture of the search kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_search -s 4 -c 2 -f -o $OUT/prof_search_$TAG \
    python bench.py --nb $NB --steps 3 --warmup 3 --L_pq $L --no_cpu_baseline > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
