#!/bin/bash
# Round-2 GPU call C (2 GPUs): single-process multi-GPU C-ABI + sharded CLIs, the fixed kNN tests, the 1M build-quality test,
# the bipartite kernel on the flattened lists, then the torchrun bench line at N = 2 (reference arm first, like the driver).
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpu_r2c.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_group.py tests/test_host_cli.py tests/test_gpu_knn.py tests/test_gpu_build.py tests/test_gpu_search.py -q -m gpu > $OUT/pytest_r2c.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_r2c.log
tail -8 $OUT/pytest_r2c.log
timeout 600 python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000 --out $OUT/bipartite_1M3M_r2c.json > /dev/null 2> $OUT/bipartite_1M3M_r2c.err; echo "bipartite rc=$?"
grep "L_pq" $OUT/bipartite_1M3M_r2c.err | cut -c1-250
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $OUT/bench_r2c_ref_n2.json 2> $OUT/bench_r2c_ref_n2.err; echo "ref n2 rc=$?"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_r2c_n2.json 2> $OUT/bench_r2c_n2.err; echo "bench n2 rc=$?"
grep -E "sub-record|failed|Error|error|L_pq=" $OUT/bench_r2c_n2.err | head -40
ls -la $OUT | tail -8
