#!/bin/bash
# Round-2 GPU call T (8 GPUs): the final tree — `torchrun ... bench.py --gpus 8` with every sub-record, the reference arm launched
# the same way, the NCCL sharded-kNN test (world 2 / 4) and the C-ABI group tests.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_r2t_n8.json 2> $OUT/bench_r2t_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2t_n8.json').read().strip().splitlines()[-1])
print('headline', l['value'], l['e2e']['value'], l['roofline']['frac'], l['operating_point']['L_pq'], l['operating_point']['recall'], l.get('clocks'))
for k in ('knn','webvid','replicas','sharded_parity'):
    r=l.get(k,{}); print(k, r.get('value'), r.get('roofline',{}).get('frac'), r.get('seconds'), r.get('phases_ms_one_batch'), r.get('ids_identical_to_reference_merge'), r.get('error'))
PY
timeout 900 python -m pytest tests/test_gpu_sharded_nccl.py tests/test_gpu_group.py -q -m gpu > $OUT/pytest_r2t.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_r2t.log
timeout 600 $TR --master-port 29521 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $OUT/bench_r2t_n8_reference.json 2> $OUT/bench_r2t_n8_reference.err; echo "reference arm rc=$?"; tail -c 600 $OUT/bench_r2t_n8_reference.json
