#!/bin/bash
# Round-2 GPU call M: query ordering (tickets hand the batch out grouped by nearest landmark) — parity + A/B on 10 k and 100 k batches.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_fullsize.py -q -m gpu > $OUT/pytest_r2m.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r2m.log
for ORD in 1 0; do
  RG_QUERY_ORDER=$ORD timeout 900 python bench.py --steps 10 --warmup 3 --extras batch --no_cpu_baseline > $OUT/bench_r2m_ord$ORD.json 2> $OUT/bench_r2m_ord$ORD.err; echo "bench order=$ORD rc=$?"
  python - <<PY
import json
l=json.loads(open('gpurun_out/bench_r2m_ord$ORD.json').read().strip().splitlines()[-1])
print('order=$ORD 10k:', round(l['value']), 'frac', round(l['roofline']['frac'],4), 'kernel_ms', round(l['roofline']['kernel_ms'],3), 'e2e', round(l['e2e']['value']), '| 100k:', round(l['batch100k']['value']), 'frac', round(l['batch100k']['roofline']['frac'],4))
PY
done
