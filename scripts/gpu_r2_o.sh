#!/bin/bash
# Round-2 GPU call O: chunk schedule under a shared bound, one rank's share of the 8-way sharded kNN on one GPU.
OUT=gpurun_out
mkdir -p $OUT
for G in 2 4; do RG_KNN_EXT_GROWTH=$G timeout 300 python scripts/bench_knn_shard.py 1250000 8 2>/dev/null | tee -a $OUT/knn_shard_schedules_r2o.txt; done
RG_KNN_EXT_GROWTH=2 RG_KNN_WHOLE_SHARD=1 timeout 300 python scripts/bench_knn_shard.py 1250000 8 2>/dev/null | tee -a $OUT/knn_shard_schedules_r2o.txt
