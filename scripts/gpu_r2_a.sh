#!/bin/bash
# Round-2 GPU call A: parity suites touched this round (A2 pin, kNN seed / MKL at scale), kNN A/B (seed on/off, per-phase trace),
# bipartite sweep at two training-query ratios + one ncu full capture of the bipartite kernel.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_r2a.txt 2>&1
nproc >> $OUT/gpu_r2a.txt
timeout 1500 python -m pytest tests/test_gpu_search.py tests/test_gpu_knn.py tests/test_host_cli.py -q -m gpu -x > $OUT/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_r2a.log
tail -4 $OUT/pytest_r2a.log
for seed in 1 0; do
  RG_KNN_SEED=$seed timeout 600 python scripts/bench_knn.py 10000000 151552 200 100 tc > $OUT/knn_r2a_seed$seed.json 2> $OUT/knn_r2a_seed$seed.err; echo "knn seed=$seed rc=$?"; cat $OUT/knn_r2a_seed$seed.json
  RG_KNN_SEED=$seed timeout 600 python scripts/bench_knn.py 1250000 151552 200 100 tc > $OUT/knn_r2a_shard_seed$seed.json 2> $OUT/knn_r2a_shard_seed$seed.err; cat $OUT/knn_r2a_shard_seed$seed.json
done
RG_KNN_TRACE=1 timeout 600 python scripts/bench_knn.py 10000000 151552 200 100 tc > $OUT/knn_r2a_trace.json 2> $OUT/knn_r2a_trace.err; cat $OUT/knn_r2a_trace.json
RG_KNN_TRACE=1 timeout 600 python scripts/bench_knn.py 1250000 151552 200 100 tc > $OUT/knn_r2a_shard_trace.json 2> $OUT/knn_r2a_shard_trace.err; cat $OUT/knn_r2a_shard_trace.json
timeout 900 python scripts/bench_bipartite.py --nb 10000000 --out $OUT/bipartite_10M_r2a.json > /dev/null 2> $OUT/bipartite_10M_r2a.err; echo "bipartite 10M rc=$?"; tail -9 $OUT/bipartite_10M_r2a.err
timeout 900 python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000 --out $OUT/bipartite_1M3M_r2a.json > /dev/null 2> $OUT/bipartite_1M3M_r2a.err; echo "bipartite 1M/3M rc=$?"; tail -9 $OUT/bipartite_1M3M_r2a.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_search -s 10 -c 1 -f -o $OUT/prof_bipartite_r2a \
    python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000 > $OUT/ncu_bip_r2a.log 2>&1
ls -la $OUT | tail -15
