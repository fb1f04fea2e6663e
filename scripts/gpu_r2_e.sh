#!/bin/bash
# Round-2 GPU call E: the phase-1 replay of the build against the reference's 1M statistics (with per-phase degree trace), host
# CLI tests on the streaming loader, kNN default schedule check.
OUT=gpurun_out
mkdir -p $OUT
RG_BUILD_TRACE=1 timeout 900 python -m pytest tests/test_gpu_build.py -q -m gpu -s > $OUT/pytest_r2e_build.log 2>&1; echo "pytest build rc=$?"
grep -E "rg build|passed|failed|AssertionError" $OUT/pytest_r2e_build.log | tail -30; head -2 $OUT/build_1M_vs_reference.csv
timeout 900 python -m pytest tests/test_host_cli.py tests/test_gpu_group.py -q -m gpu > $OUT/pytest_r2e_cli.log 2>&1; echo "pytest cli rc=$?"; tail -4 $OUT/pytest_r2e_cli.log
for NB in 1250000 10000000; do
  RG_KNN_TRACE=1 timeout 600 python scripts/bench_knn.py $NB 151552 200 100 tc 2>/dev/null | tee -a $OUT/knn_r2e.json | cut -c1-400
done
