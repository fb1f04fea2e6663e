#!/bin/bash
# Round-2 GPU call Q (1 GPU): bound re-sync inside the shard — exactness test and the one-rank microbenchmark.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_knn.py -q -m gpu -k "resync or shared_tau" > $OUT/pytest_r2q.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_r2q.log
timeout 300 python scripts/bench_knn_shard.py 1250000 8 2>/dev/null | tee $OUT/knn_shard_resync_r2q.txt
