#!/bin/bash
# parity tests of the search path under the row-ring / visited-set variants, then kernel-only timings over the launch knobs
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
for cfg in "8 tags" "8 hash" "16 hash"; do
  set -- $cfg
  RG_ROWS_PER_BATCH=$1 RG_VISITED=$2 timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_build.py -x -q -m gpu > $OUT/pytest_search_R$1_$2_$TAG.log 2>&1
  echo "R=$1 visited=$2: $(tail -n 1 $OUT/pytest_search_R$1_$2_$TAG.log)"
done
timeout 1500 python scripts/tune_search.py 10000000 130 > $OUT/tune_$TAG.csv 2> $OUT/tune_$TAG.err; tail -n 3 $OUT/tune_$TAG.err; cat $OUT/tune_$TAG.csv
