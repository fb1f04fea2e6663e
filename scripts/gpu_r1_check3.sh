#!/bin/bash
# One gpurun call (1 GPU): search + host-CLI GPU tests, then the L_pq sweep with the merge insert (RG_QUEUE_MERGE=1, default)
# for the single- and double-buffered row ring; compare with sweep_R8.csv / sweep_R16.csv of gpu_r1_check2.sh (merge off).
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_search.py tests/test_host_cli.py -x -q -m gpu > $OUT/pytest_search.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_search.log
timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_R8_merge.csv --no_knn --no_cpu_baseline > $OUT/bench_R8_merge.json 2> $OUT/bench_R8_merge.err; echo "R8 rc=$?"
RG_ROWS_PER_BATCH=16 timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_R16_merge.csv --no_knn --no_cpu_baseline > $OUT/bench_R16_merge.json 2> $OUT/bench_R16_merge.err; echo "R16 rc=$?"
RG_QUEUE_MERGE=0 timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_R8_nomerge.csv --no_knn --no_cpu_baseline > $OUT/bench_R8_nomerge.json 2> $OUT/bench_R8_nomerge.err; echo "R8 nomerge rc=$?"
