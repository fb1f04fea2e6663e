#!/bin/bash
# Round-2 GPU call D: the rebuilt reverse-edge phases against the reference's 1M statistics, kNN chunk-size A/B (re-rank rows in
# L2), the fixed kNN parity tests, a 10M bench line with the kNN / batch sub-records.
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_build.py tests/test_gpu_knn.py tests/test_gpu_search.py -q -m gpu > $OUT/pytest_r2d.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_r2d.log
tail -6 $OUT/pytest_r2d.log; head -2 $OUT/build_1M_vs_reference.csv
for CM in 0 65536 262144 4194304; do
  for NB in 1250000 10000000; do
    if [ $CM = 0 ]; then unset RG_KNN_CHUNK_MAX; else export RG_KNN_CHUNK_MAX=$CM; fi
    RG_KNN_TRACE=1 timeout 600 python scripts/bench_knn.py $NB 151552 200 100 tc 2> $OUT/knn_r2d_$CM_$NB.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); p=d.get('phases_ms_of_last_call') or {}
    print('chunk_max=$CM nb=$NB', 'frac %.3f resident %.3f' % (d['frac_of_measured_bf16_sustained'], d['resident_base']['frac_of_measured_bf16_sustained']), {k: round(v,1) for k,v in p.items()})
" | tee -a $OUT/knn_chunk_ab_r2d.txt
  done
done
unset RG_KNN_CHUNK_MAX
timeout 1500 python bench.py --steps 20 --warmup 5 --extras batch,knn,bipartite > $OUT/bench_r2d.json 2> $OUT/bench_r2d.err; echo "bench rc=$?"
grep -E "sub-record|failed|Error|error|BuildRoarGraph|L_pq=" $OUT/bench_r2d.err | head -30
ls -la $OUT | tail -6
