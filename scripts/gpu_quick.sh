#!/bin/bash
# Short GPU-box visit: parity tests (optional), bench of the target and of C2's grid, launch list.
# Usage: bash scripts/gpu_quick.sh <tag> [tests: all|fast|none] [ncu: 0|1]
TAG=${1:-quick}; TESTS=${2:-fast}; NCU=${3:-0}
export TCFD_CHUNK_MB=${TCFD_CHUNK_MB:-100000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$TESTS" = "all" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
elif [ "$TESTS" = "fast" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "not drift and not 1024" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
fi
[ -f $OUT/pytest_gpu.log ] && tail -4 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_512.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --n 256 --steps 60 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_256.json 2>> $OUT/bench.err
python - <<PY
import json
for f in ("bench_512","bench_256"):
    try:
        d=json.load(open("$OUT/%s.json"%f))
        print(f, "steps/s=%.1f"%d["value"], "frac=%.3f"%d["roofline"]["frac"], "e2e=", d.get("e2e") and round(d["e2e"]["value"],1), {k:(round(v["us_per_launch"],1) if isinstance(v,dict) and v["us_per_launch"] else None) for k,v in d["kernels"].items() if isinstance(v,dict)})
    except Exception as e:
        print(f, "failed", e)
PY
tail -5 $OUT/bench.err
if [ "$NCU" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 33 -c 44 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns2d_ -s 36 -c 2 -o $OUT/prof_full -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
  ls -la $OUT
fi
