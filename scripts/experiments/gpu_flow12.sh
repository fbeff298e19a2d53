#!/bin/bash
OUT=gpurun_out/flow12
mkdir -p $OUT
for G in 1,1,4 3,4,4 2,2,4; do
  echo "== 512 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 100 --configs 1:64:64,1:64:4,1:64:6,1:64:8,1:64:10,1:64:12,1:64:16,1:64:24 2> $OUT/s512_$G.err | tee $OUT/s512_$G.jsonl | cut -c1-125
  tail -1 $OUT/s512_$G.err
done
echo "== 256"
TCFD_FLOW_G=1,1,8 timeout 300 python scripts/sweep_flow.py --n 256 --batch 64 --steps 200 --configs 1:64:64,1:64:8,1:64:16,1:64:32 2> $OUT/s256.err | tee $OUT/s256.jsonl | cut -c1-125
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
for WH in 8 12; do
TCFD_FLOW_G=1,1,4 timeout 300 ncu --metrics $M --clock-control none -k regex:ns2d_flow -s 3 -c 1 --csv --log-file $OUT/m_$WH.csv python scripts/sweep_flow.py --n 512 --batch 64 --steps 1 --configs 1:64:$WH > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/m_$WH.csv")) if len(r)>10]
h=rows[0]
print("Wh=$WH", {r[h.index("Metric Name")]: r[h.index("Metric Value")] for r in rows[1:]})
PY
done
