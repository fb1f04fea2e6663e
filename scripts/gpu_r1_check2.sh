#!/bin/bash
# One gpurun call (1 GPU): the GPU half of tests/test_host_cli.py (re-hosted + drop-in CLIs), the L_pq sweep with the
# single-buffered (R=8) and double-buffered (R=16) row ring, and the reference arm of bench.py.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_host_cli.py -x -q -m gpu > $OUT/pytest_cli.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_cli.log
timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_R8.csv --no_knn --no_cpu_baseline > $OUT/bench_R8.json 2> $OUT/bench_R8.err; echo "R8 rc=$?"
RG_ROWS_PER_BATCH=16 timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_R16.csv --no_knn --no_cpu_baseline > $OUT/bench_R16.json 2> $OUT/bench_R16.err; echo "R16 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -3 $OUT/bench_ref.err | cut -c1-300; cat $OUT/bench_ref.json | cut -c1-1500
