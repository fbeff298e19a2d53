#!/bin/bash
OUT=gpurun_out/flow10
mkdir -p $OUT
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum
for cfg in "1 8 1,1,4" "0 8 1,1,4" "1 6 1,1,4" "0 12 1,1,4" "0 64 1,1,4" "1 8 3,4,-29"; do
  set -- $cfg
  echo "== persist=$1 W=$2 G=$3"
  TCFD_FLOW_PERSIST=$1 TCFD_FLOW_G=$3 timeout 300 ncu --metrics $M --clock-control none -k regex:ns2d_flow -s 3 -c 1 --csv --log-file $OUT/m_$1_$2_$3.csv \
    python scripts/sweep_flow.py --n 512 --batch 64 --steps 1 --configs 1:$2 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/m_$1_$2_$3.csv")) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    print("  ", r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
done
