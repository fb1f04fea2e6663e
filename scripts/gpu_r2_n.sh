#!/bin/bash
# Round-2 GPU call N (8 GPUs): the 10 M x 10 M kNN record after the whole-shard schedule; C-ABI group tests at G = 2 / 4.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --extras knn > $OUT/bench_r2n_n8.json 2> $OUT/bench_r2n_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2n_n8.json').read().strip().splitlines()[-1])
k=l['knn']; print('knn n8', k.get('value'), k.get('seconds'), k.get('roofline',{}).get('frac'), k.get('phases_ms_one_batch'), k.get('check'), k.get('error'))
print('headline', l['value'], l['operating_point']['L_pq'])
PY
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_host_cli.py -q -m gpu -k "group or sharded" > $OUT/pytest_r2n.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r2n.log
