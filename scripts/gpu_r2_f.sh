#!/bin/bash
# Round-2 GPU call F: kNN two-lane A/B + parity tests, then the ncu evidence of the round: launch list of the default bench
# command, full captures of the search kernel at the headline L_pq and at L_pq = 2000, of the bipartite kernel, of the laion
# (d = 512) search, and of the kNN filter / merge kernels.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_knn.py -q -m gpu > $OUT/pytest_r2f_knn.log 2>&1; echo "pytest knn rc=$?"; tail -3 $OUT/pytest_r2f_knn.log
for LANES in 2 1; do for NB in 1250000 10000000; do
  RG_KNN_LANES=$LANES timeout 600 python scripts/bench_knn.py $NB 151552 200 100 tc 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('lanes=$LANES nb=$NB frac %.3f resident %.3f (%.1f ms)' % (d['frac_of_measured_bf16_sustained'], d['resident_base']['frac_of_measured_bf16_sustained'], d['resident_base']['search_s']*1e3))
" | tee -a $OUT/knn_lanes_ab_r2f.txt
done; done
# populate the graph cache + the bench line of this build
timeout 1500 python bench.py --steps 20 --warmup 5 --extras batch,knn,bipartite > $OUT/bench_r2f.json 2> $OUT/bench_r2f.err; echo "bench rc=$?"
grep -E "sub-record|failed|Error|BuildRoarGraph|L_pq=" $OUT/bench_r2f.err | head -20
L=$(python -c "import json;print(json.loads(open('$OUT/bench_r2f.json').read().strip().splitlines()[-1])['operating_point']['L_pq'])" 2>/dev/null || echo 120)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2f.csv \
    python bench.py --steps 2 --warmup 1 --extras none --no_cpu_baseline > $OUT/ncu_list_r2f.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_search -s 6 -c 1 -f -o $OUT/prof_search_L${L}_r2f \
    python bench.py --steps 3 --warmup 3 --L_pq $L --extras none --no_cpu_baseline > $OUT/ncu_a_r2f.log 2>&1; echo "ncu search rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_search -s 4 -c 1 -f -o $OUT/prof_search_L2000_r2f \
    python bench.py --steps 3 --warmup 3 --L_pq 2000 --extras none --no_cpu_baseline > $OUT/ncu_b_r2f.log 2>&1; echo "ncu L2000 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_search -s 8 -c 1 -f -o $OUT/prof_bipartite_r2f \
    python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000 > $OUT/ncu_c_r2f.log 2>&1; echo "ncu bipartite rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_tc_filter|knn_merge_fast' -s 24 -c 4 -f -o $OUT/prof_knn_r2f \
    python scripts/bench_knn.py 10000000 151552 200 100 tc > $OUT/ncu_d_r2f.log 2>&1; echo "ncu knn rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:beam_search -s 6 -c 1 -f -o $OUT/prof_search_laion_r2f \
    python bench.py --nb 10000000 --dim 512 --dist cosine --steps 3 --warmup 3 --L_pq 120 --extras none --no_cpu_baseline > $OUT/ncu_e_r2f.log 2>&1; echo "ncu laion rc=$?"
ls -la $OUT | tail -12
