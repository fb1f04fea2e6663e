#!/bin/bash
# Round-2 GPU call P (8 GPUs): 10 M x 10 M kNN with the bound's all-reduce issued one batch ahead and the all-to-all slice
# exchange; the NCCL sharded-kNN test at world 2 / 4.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --extras knn > $OUT/bench_r2p_n8.json 2> $OUT/bench_r2p_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2p_n8.json').read().strip().splitlines()[-1])
k=l['knn']; print('knn n8', k.get('value'), k.get('seconds'), k.get('roofline',{}).get('frac'), k.get('phases_ms_one_batch'), k.get('check'), k.get('error'))
print('headline', l['value'], l['operating_point']['L_pq'])
PY
timeout 900 python -m pytest tests/test_gpu_sharded_nccl.py -q -m gpu > $OUT/pytest_r2p.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_r2p.log
