#!/bin/bash
# Round-2 GPU call K (8 GPUs): the driver's scaling command at N = 8 — torchrun bench.py --gpus 8 with every sub-record
# (100 k-query batch base-sharded headline, sharded id parity, the whole 10 M x 10 M kNN with the shared tau, webvid, replicas).
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu_r2k.txt 2>&1
RG_BUILD_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_r2k_n8.json 2> $OUT/bench_r2k_n8.err; echo "bench n8 rc=$?"
grep -E "sub-record|failed|Error|error|L_pq=" $OUT/bench_r2k_n8.err | sort | uniq -c | sort -rn | head -30
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_r2k_n8.json').read().strip().splitlines()[-1])
print('headline', l['value'], l['e2e']['value'], l['operating_point']['L_pq'], l['operating_point']['recall'], l['roofline']['frac'], l.get('sharded_parity'))
for key in ('batch10k','knn','webvid','replicas'):
    r=l.get(key,{})
    print(key, r.get('value'), r.get('L_pq'), r.get('recall'), (r.get('roofline') or {}).get('frac'), r.get('seconds'), r.get('error'), r.get('phases_ms_one_batch'), r.get('check'))
PY
