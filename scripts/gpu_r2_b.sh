#!/bin/bash
# Round-2 GPU call B: the whole GPU suite, a shake-down of bench.py with every sub-record at 1/10 size, the full default
# bench line (N = 1), and the bipartite sweep A/B (row ring 8 vs 16).
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_r2b.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_r2b.log
tail -6 $OUT/pytest_r2b.log
timeout 900 python bench.py --nb 1000000 --extras_div 10 --steps 5 --warmup 3 > $OUT/bench_r2b_small.json 2> $OUT/bench_r2b_small.err; echo "bench small rc=$?"
grep -E "sub-record|failed|Error|error" $OUT/bench_r2b_small.err | head -30
timeout 1500 python bench.py --steps 20 --warmup 5 --sweep $OUT/sweep_r2b.csv > $OUT/bench_r2b.json 2> $OUT/bench_r2b.err; echo "bench full rc=$?"
grep -E "sub-record|failed|Error|error" $OUT/bench_r2b.err | head -30
for R in 8 16; do
  RG_ROWS_PER_BATCH=$R timeout 600 python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000 --out $OUT/bipartite_1M3M_r2b_R$R.json > /dev/null 2> $OUT/bipartite_1M3M_r2b_R$R.err; echo "bipartite R=$R rc=$?"
  grep "L_pq" $OUT/bipartite_1M3M_r2b_R$R.err | cut -c1-260
done
ls -la $OUT | tail -12
