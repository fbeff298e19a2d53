#!/bin/bash
# chunk-size sweep of the L2-resident schedule (TCFD_CHUNK_MB) on the target workload
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
for mb in 10 20 40 60 80 100000; do
  TCFD_CHUNK_MB=$mb timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_chunk_$mb.json 2>> $OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_chunk_$mb.json"))
print("chunk_mb=$mb", "steps/s=%.1f"%d["value"], "frac=%.3f"%d["roofline"]["frac"], {k:(round(v["us_per_launch"],1) if isinstance(v,dict) and v["us_per_launch"] else None) for k,v in d["kernels"].items() if isinstance(v,dict)})
PY
done
