#!/bin/bash
# Round-2 GPU call I (2 GPUs): the shared-tau kNN protocol (one GPU test, C-ABI group at G = 2, torchrun bench kNN record at N = 2).
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_group.py -q -m gpu -k "shared_tau or group" > $OUT/pytest_r2i.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r2i.log
timeout 600 python -m pytest tests/test_host_cli.py -q -m gpu -k sharded >> $OUT/pytest_r2i.log 2>&1; echo "pytest cli rc=$?"; tail -2 $OUT/pytest_r2i.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --extras knn,sharded_parity > $OUT/bench_r2i_n2.json 2> $OUT/bench_r2i_n2.err; echo "bench n2 rc=$?"
grep -E "sub-record|failed|Error|error|L_pq=" $OUT/bench_r2i_n2.err | head -20
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2i_n2.json').read().strip().splitlines()[-1])
k=l['knn']; print('knn n2', k.get('value'), k.get('seconds'), k.get('roofline',{}).get('frac'), k.get('phases_ms_one_batch'), k.get('check'), k.get('error'))
print('headline', l['value'], l['operating_point']['L_pq'], l.get('sharded_parity'))
PY
