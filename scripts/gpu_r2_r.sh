#!/bin/bash
# Round-2 GPU call R (1 GPU): merge-by-rank list update — the whole kNN test file, then A/B on one rank's shard and on 10 M rows.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_knn.py -q -m gpu -x > $OUT/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_r2r.log
for m in 1 0; do RG_KNN_MERGE_RANK=$m timeout 300 python scripts/bench_knn_shard.py 1250000 8 2>/dev/null | tee -a $OUT/knn_merge_rank_r2r.txt; done
for m in 1 0; do RG_KNN_MERGE_RANK=$m timeout 300 python scripts/bench_knn_shard.py 10000000 1 2>/dev/null | tee -a $OUT/knn_merge_rank_r2r.txt; done
