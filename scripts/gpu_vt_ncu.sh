#!/bin/bash
# DRAM traffic of the search kernel with the tag arrays vs the hashed visited table (same launch geometry)
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
export RG_CTAS_PER_SM=${CTAS:-4}
python bench.py --nb 10000000 --steps 3 --warmup 3 --L_pq 130 --no_cpu_baseline --no_knn > /dev/null 2>&1   # builds the graph cache
for v in tags hash; do
  RG_VISITED=$v RG_VTAB_MB=512 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_atom.sum \
      --clock-control none -k regex:beam_search -s 3 -c 2 --csv --log-file $OUT/vt_ncu_${v}_$TAG.csv \
      python bench.py --nb 10000000 --steps 3 --warmup 3 --L_pq 130 --no_cpu_baseline --no_knn > $OUT/vt_ncu_${v}_$TAG.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("$OUT/vt_ncu_${v}_$TAG.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
h=rows[hi]
for r in rows[hi+1:]:
    if len(r)>=len(h) and r[0]=="0": print("$v", r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
done
