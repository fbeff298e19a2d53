#!/bin/bash
OUT=gpurun_out/sconv1
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sconv -s 60 -c 40 --csv --log-file $OUT/launches.csv \
  python scripts/bench_sconv.py --iters 4 > $OUT/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/sconv1/launches.csv")) if len(r)>10]
h=rows[0]
agg=collections.OrderedDict()
for r in rows[1:]:
    k=(r[h.index("ID")], r[h.index("Kernel Name")][:60], r[h.index("Grid Size")], r[h.index("Block Size")])
    agg.setdefault(k,{})[r[h.index("Metric Name")]]=r[h.index("Metric Value")]
for k,v in list(agg.items())[:30]:
    print(k[0],k[1],k[2],k[3], v.get("gpu__time_duration.sum"), v.get("dram__bytes_read.sum"), v.get("dram__bytes_write.sum"))
PY
