#!/bin/bash
# Round-2 GPU call S (1 GPU): the whole GPU test suite, smoke(), both bench arms with default flags, and a fresh ncu capture of
# the kNN filter + re-rank kernels — the final tree.
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
timeout 1500 python -m pytest tests -q -m gpu -x > $OUT/pytest_r2s.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_r2s.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_r2s.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_r2s.log
timeout 900 python bench.py --impl reference > $OUT/bench_r2s_reference.json 2> $OUT/bench_r2s_reference.err; echo "reference arm rc=$?"
timeout 1500 python bench.py > $OUT/bench_r2s.json 2> $OUT/bench_r2s.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2s.json').read().strip().splitlines()[-1])
print('headline', l['value'], l['e2e']['value'], l['roofline']['frac'], l['operating_point'], l['cpu_baseline']['value'], l.get('clocks'))
for k in ('batch100k','knn','bipartite','laion'):
    r=l.get(k,{}); print(k, r.get('value'), r.get('roofline',{}).get('frac'), r.get('seconds'), r.get('error'))
r=json.loads(open('gpurun_out/bench_r2s_reference.json').read().strip().splitlines()[-1]); print('reference', r.get('value'), r.get('cpu_baseline',{}).get('cores'))
PY
cap() {  # name, skip, count, kernel regex (demangled), command...
  local name=$1 skip=$2 cnt=$3 kre=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kre" -s $skip -c $cnt -f -o /tmp/ncu/$name "$@" > /tmp/ncu/$name.log 2>&1
  echo "ncu $name rc=$?"; tail -2 /tmp/ncu/$name.log
  python scripts/ncu_summary.py /tmp/ncu/$name.ncu-rep 25 > $OUT/ncu_$name.txt 2>&1
}
cap knn_10M_r2s 12 4 'knn_tc_filter|knn_merge_fast' python scripts/bench_knn.py 10000000 151552 200 100 tc
du -sh $OUT
