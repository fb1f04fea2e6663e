#!/bin/bash
# One gpurun call (1 GPU): the whole GPU test-suite (no -x), then the L_pq sweep with the shared-memory visited table
# (RG_VISITED=smem) against the default tag arrays on the same box.
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_full.log
timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_tags.csv --no_knn --no_cpu_baseline > $OUT/bench_tags.json 2> $OUT/bench_tags.err; echo "tags rc=$?"
RG_VISITED=smem timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_smem.csv --no_knn --no_cpu_baseline > $OUT/bench_smem.json 2> $OUT/bench_smem.err; echo "smem rc=$?"
RG_VISITED=smem RG_ROWS_PER_BATCH=8 timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_smem_R8.csv --no_knn --no_cpu_baseline > $OUT/bench_smem_R8.json 2> $OUT/bench_smem_R8.err; echo "smem R8 rc=$?"
RG_VISITED=smem RG_ROWS_PER_BATCH=16 timeout 600 python bench.py --steps 10 --warmup 3 --sweep $OUT/sweep_smem_R16.csv --no_knn --no_cpu_baseline > $OUT/bench_smem_R16.json 2> $OUT/bench_smem_R16.err; echo "smem R16 rc=$?"
tail -2 $OUT/bench_smem.err | cut -c1-300
