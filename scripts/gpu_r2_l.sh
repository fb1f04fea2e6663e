#!/bin/bash
# Round-2 GPU call L: kNN parity tests after the whole-shard schedule, ring-depth A/B at long queues.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_group.py -q -m gpu > $OUT/pytest_r2l.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r2l.log
for R in 8 16; do
  RG_ROWS_PER_BATCH=$R timeout 900 python bench.py --steps 3 --warmup 3 --extras none --no_cpu_baseline --L_pq 110 --sweep $OUT/sweep_r2l_R$R.csv > /dev/null 2> $OUT/bench_r2l_R$R.err; echo "sweep R=$R rc=$?"
done
python - <<'PY'
import csv
a={r['L_pq']:r for r in csv.DictReader(open('gpurun_out/sweep_r2l_R8.csv'))}
b={r['L_pq']:r for r in csv.DictReader(open('gpurun_out/sweep_r2l_R16.csv'))}
for L in a:
    if int(float(L)) >= 500: print('L',L,'R8 frac',a[L]['hbm_frac'],'| R16 frac',b[L]['hbm_frac'])
PY
