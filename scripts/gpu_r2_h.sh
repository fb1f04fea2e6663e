#!/bin/bash
# Round-2 GPU call H: the default bench line with every sub-record (N = 1 evidence), the launch list of the timed step's
# kernels, and targeted ncu captures (bipartite kernel, laion d = 512 query search, single-lane kNN filter + merge), summarised
# on the box.
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
timeout 2400 python bench.py --steps 20 --warmup 5 > $OUT/bench_r2h.json 2> $OUT/bench_r2h.err; echo "bench rc=$?"
grep -E "sub-record|failed|Error|BuildRoarGraph|L_pq=" $OUT/bench_r2h.err | head -30
L=$(python -c "import json;print(json.loads(open('$OUT/bench_r2h.json').read().strip().splitlines()[-1])['operating_point']['L_pq'])" 2>/dev/null || echo 110)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'beam_search|tags_|topk_' -c 60 --csv \
    --log-file $OUT/launches_r2h.csv python bench.py --steps 5 --warmup 3 --L_pq $L --extras none --no_cpu_baseline > /tmp/ncu/list.log 2>&1; echo "launch list rc=$?"
cap() {  # name, skip, count, kernel regex (demangled), command...
  local name=$1 skip=$2 cnt=$3 kre=$4; shift 4
  timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kre" -s $skip -c $cnt -f -o /tmp/ncu/$name "$@" > /tmp/ncu/$name.log 2>&1
  echo "ncu $name rc=$?"; tail -2 /tmp/ncu/$name.log
  python scripts/ncu_summary.py /tmp/ncu/$name.ncu-rep 25 > $OUT/ncu_$name.txt 2>&1
}
# bench_bipartite.py launches per L_pq: warm-up (100 queries), 3 timed; L sweep 10, 20, 50 -> launch 9..11 are L_pq = 50
cap bipartite_1M3M_L50 9 1 'beam_search_kernel<0, 1' python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000
cap knn_10M 12 4 'knn_tc_filter|knn_merge_fast' python scripts/bench_knn.py 10000000 151552 200 100 tc
# laion: mode-0 launches only (the build's own searches are mode 2): recall pass, 3 warm-ups, then the timed steps
cap search_laion_10M_L120 5 1 'beam_search_kernel<0, 0' python bench.py --nb 10000000 --dim 512 --dist cosine --steps 3 --warmup 3 --L_pq 120 --extras none --no_cpu_baseline
ls -la $OUT | tail -10; du -sh $OUT
