#!/bin/bash
# Round-2 GPU call G: lazy queue (long L_pq) parity + sweep A/B, kNN two-lane A/B with the persistent merge grid, and the ncu
# evidence of the round summarised ON the box (reports are ~17 MB each; only text summaries travel back).
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
timeout 1500 python -m pytest tests/test_gpu_search.py tests/test_gpu_fullsize.py tests/test_gpu_build.py tests/test_gpu_knn.py -q -m gpu > $OUT/pytest_r2g.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_r2g.log
for LANES in 2 1; do for NB in 1250000 10000000; do
  RG_KNN_LANES=$LANES timeout 600 python scripts/bench_knn.py $NB 151552 200 100 tc 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('lanes=$LANES nb=$NB frac %.3f resident %.3f (%.1f ms)' % (d['frac_of_measured_bf16_sustained'], d['resident_base']['frac_of_measured_bf16_sustained'], d['resident_base']['search_s']*1e3))
" | tee -a $OUT/knn_lanes_ab_r2g.txt
done; done
timeout 1500 python bench.py --steps 20 --warmup 5 --extras batch,knn,bipartite --sweep $OUT/sweep_r2g_lazy.csv > $OUT/bench_r2g.json 2> $OUT/bench_r2g.err; echo "bench rc=$?"
grep -E "sub-record|failed|Error|BuildRoarGraph|L_pq=" $OUT/bench_r2g.err | head -20
RG_LAZY_QUEUE_MIN_L=0 timeout 900 python bench.py --steps 5 --warmup 3 --extras none --no_cpu_baseline --sweep $OUT/sweep_r2g_eager.csv > /dev/null 2> $OUT/bench_r2g_eager.err; echo "eager sweep rc=$?"
python - <<'PY'
import csv
a={r['L_pq']:r for r in csv.DictReader(open('gpurun_out/sweep_r2g_lazy.csv'))}
b={r['L_pq']:r for r in csv.DictReader(open('gpurun_out/sweep_r2g_eager.csv'))}
for L in a:
    if int(float(L)) in (100,200,260,300,400,500,700,1000,1400,2000): print('L',L,'lazy frac',a[L]['hbm_frac'],'qps',a[L]['qps'],'| eager frac',b[L]['hbm_frac'],'qps',b[L]['qps'],'| recall',a[L]['recall'],b[L]['recall'])
PY
L=$(python -c "import json;print(json.loads(open('$OUT/bench_r2g.json').read().strip().splitlines()[-1])['operating_point']['L_pq'])" 2>/dev/null || echo 110)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2g.csv \
    python bench.py --steps 2 --warmup 1 --extras none --no_cpu_baseline > /tmp/ncu/list.log 2>&1; echo "launch list rc=$?"
cap() {  # name, skip, count, kernel regex, command...
  local name=$1 skip=$2 cnt=$3 kre=$4; shift 4
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c $cnt -f -o /tmp/ncu/$name "$@" > /tmp/ncu/$name.log 2>&1
  echo "ncu $name rc=$?"
  python scripts/ncu_summary.py /tmp/ncu/$name.ncu-rep 25 > $OUT/ncu_$name.txt 2>&1
}
cap search_L${L}_10M 6 1 beam_search python bench.py --steps 3 --warmup 3 --L_pq $L --extras none --no_cpu_baseline
cap search_L2000_10M 4 1 beam_search python bench.py --steps 3 --warmup 3 --L_pq 2000 --extras none --no_cpu_baseline
cap bipartite_1M3M 8 1 beam_search python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000
cap knn_10M 24 4 'knn_tc_filter|knn_merge_fast' python scripts/bench_knn.py 10000000 151552 200 100 tc
cap search_laion_10M 6 1 beam_search python bench.py --nb 10000000 --dim 512 --dist cosine --steps 3 --warmup 3 --L_pq 120 --extras none --no_cpu_baseline
cp /tmp/ncu/search_L${L}_10M.ncu-rep $OUT/ 2>/dev/null
ls -la $OUT | tail -14; du -sh $OUT
