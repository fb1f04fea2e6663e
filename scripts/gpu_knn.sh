#!/bin/bash
# One gpurun call for the brute-force kNN path: GPU parity tests, a throughput run, the ncu launch list (with the tensor
# pipe share per launch) and one full capture of the tensor-core filter kernel on its largest chunk.
# usage: scripts/gpu_knn.sh [TAG] [NB] [NQ] [TESTS]      (run from the repo root on the GPU box)
TAG=${1:-r1}
NB=${2:-10000000}
NQ=${3:-151552}
TESTS=${4:-tests/test_gpu_knn.py}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
if [ "$TESTS" != "none" ]; then
  timeout 900 python -m pytest $TESTS -x -q -m gpu > $OUT/pytest_knn_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_knn_$TAG.log
  tail -5 $OUT/pytest_knn_$TAG.log
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 200 > $OUT/clocks_knn_$TAG.csv &
SMI=$!
timeout 600 python scripts/bench_knn.py $NB $NQ 200 100 tc > $OUT/knn_$TAG.json 2> $OUT/knn_$TAG.err; echo "bench_knn rc=$?"
if [ -n "$NQ2" ]; then timeout 600 python scripts/bench_knn.py $NB $NQ2 200 100 tc >> $OUT/knn_$TAG.json 2>> $OUT/knn_$TAG.err; fi
kill $SMI
cat $OUT/knn_$TAG.json; tail -3 $OUT/knn_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    -k regex:'knn_|to_bf16|row_sqnorm' -c 200 --csv --log-file $OUT/launches_knn_$TAG.csv \
    python scripts/bench_knn.py ${NB_NCU:-2000000} $NQ 200 100 tc > $OUT/ncu_knn_list_$TAG.log 2>&1
# full capture of the filter kernel: skip the warm-up call's launches and the small early chunks, take the big ones
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_tc_filter -s ${SKIP:-17} -c 2 -f -o $OUT/prof_knn_$TAG \
    python scripts/bench_knn.py ${NB_NCU:-2000000} $NQ 200 100 tc > $OUT/ncu_knn_full_$TAG.log 2>&1
ls -la $OUT | tail -20
