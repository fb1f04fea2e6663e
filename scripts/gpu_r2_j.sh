#!/bin/bash
# Round-2 GPU call J: targeted ncu captures that earlier calls missed (bipartite kernel, laion d = 512 query search), the
# shared-tau kNN test, lazy-queue threshold sweep around L_pq = 1000.
OUT=gpurun_out
mkdir -p $OUT /tmp/ncu
timeout 600 python -m pytest tests/test_gpu_knn.py -q -m gpu -k "shared_tau" > $OUT/pytest_r2j.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_r2j.log
cap() {  # name, skip, count, kernel regex (demangled), command...
  local name=$1 skip=$2 cnt=$3 kre=$4; shift 4
  timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kre" -s $skip -c $cnt -f -o /tmp/ncu/$name "$@" > /tmp/ncu/$name.log 2>&1
  echo "ncu $name rc=$?"; tail -2 /tmp/ncu/$name.log | cut -c1-200
  python scripts/ncu_summary.py /tmp/ncu/$name.ncu-rep 25 > $OUT/ncu_$name.txt 2>&1
}
# bench_bipartite.py launches per L_pq: warm-up (100 queries), 3 timed; L sweep 10, 20, 50 -> launches 9..11 are L_pq = 50
cap bipartite_1M3M_L50 9 1 'beam_search_kernel<\(bool\)0, \(int\)1' python scripts/bench_bipartite.py --nb 1000000 --ntrain 3000000
# laion: mode-0 launches only (the build's own searches are mode 2): recall pass, 3 warm-ups, then the timed steps
cap search_laion_10M_L120 5 1 'beam_search_kernel<\(bool\)0, \(int\)0' python bench.py --nb 10000000 --dim 512 --dist cosine --steps 3 --warmup 3 --L_pq 120 --extras none --no_cpu_baseline
ls -la $OUT | tail -6; du -sh $OUT
