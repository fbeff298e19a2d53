#!/bin/bash
OUT=gpurun_out/sconv2
mkdir -p $OUT
timeout 600 python -m pytest tests/test_sconv_gpu.py -m gpu -x -q > $OUT/pytest_sconv.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sconv.log
echo "== v2 planes"
timeout 300 python scripts/bench_sconv.py 2> $OUT/bench_v2.err | tee $OUT/bench_v2.jsonl | cut -c1-330
echo "== v1 planes"
TCFD_SCONV_PLANES=1 timeout 300 python scripts/bench_sconv.py 2> $OUT/bench_v1.err | tee $OUT/bench_v1.jsonl | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sconv -s 60 -c 12 --csv --log-file $OUT/launches.csv \
  python scripts/bench_sconv.py --iters 4 > $OUT/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/sconv2/launches.csv")) if len(r)>10]
h=rows[0]
agg=collections.OrderedDict()
for r in rows[1:]:
    k=(r[h.index("ID")], r[h.index("Kernel Name")][:60], r[h.index("Grid Size")], r[h.index("Block Size")])
    agg.setdefault(k,{})[r[h.index("Metric Name")]]=r[h.index("Metric Value")]
for k,v in list(agg.items())[:12]:
    print(k[0],k[1],k[2],k[3], v.get("gpu__time_duration.sum"), v.get("dram__bytes_read.sum"), v.get("dram__bytes_write.sum"))
PY
timeout 300 python scripts/bench_fno3d.py 2>>$OUT/bench.err | tee $OUT/bench_fno3d.json
