#!/bin/bash
# parity tests of the search path under both row-ring variants, then kernel-only timings over the launch knobs
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_search.py -x -q -m gpu > $OUT/pytest_search_R8_$TAG.log 2>&1; tail -2 $OUT/pytest_search_R8_$TAG.log
RG_ROWS_PER_BATCH=16 timeout 900 python -m pytest tests/test_gpu_search.py -x -q -m gpu > $OUT/pytest_search_R16_$TAG.log 2>&1; tail -2 $OUT/pytest_search_R16_$TAG.log
timeout 1500 python scripts/tune_search.py 10000000 130 50 > $OUT/tune_$TAG.csv 2> $OUT/tune_$TAG.err; tail -3 $OUT/tune_$TAG.err; cat $OUT/tune_$TAG.csv
