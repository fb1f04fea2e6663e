#!/bin/bash
# Final round-1 evidence run (1 GPU): scripts/gpu_check.sh (whole GPU suite, smoke, 10M bench + sweep, ncu launch list and
# full capture) followed by one A/B of RG_WAVE_FIT on the cached index.
bash scripts/gpu_check.sh 10000000 r1w
RG_WAVE_FIT=1 timeout 300 python bench.py --steps 20 --warmup 3 --L_pq 130 --no_knn --no_cpu_baseline > gpurun_out/bench_wavefit.json 2> gpurun_out/bench_wavefit.err; echo "wavefit rc=$?"
python -c "
import json
for f in ('bench_r1w','bench_wavefit'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['roofline']['frac'], d['config']['launch'], d['e2e']['value'])"
