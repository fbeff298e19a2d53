#!/bin/bash
# One GPU-box visit (gpurun): parity tests, headline bench, schedule sweep with the -DTCFD_FLOW_VARIANTS
# library, region attribution, ncu captures.  Usage: bash scripts/gpu_session.sh <tag> [tests: all|fast|none]
TAG=${1:-s}; TESTS=${2:-all}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ "$TESTS" = "all" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
elif [ "$TESTS" = "fast" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "not drift and not 1024 and not horizon" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
fi
[ -f $OUT/pytest_gpu.log ] && tail -6 $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_512.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_512.json"))
    print("bench steps/s=%.1f"%d["value"], "frac=%.3f"%d["roofline"]["frac"], "e2e=", d.get("e2e") and round(d["e2e"]["value"],1))
except Exception as e:
    print("bench failed", e)
PY
if [ -n "$SWEEP" ]; then
  TCFD_LIB=$PWD/torch-cfd_b200/libtcfd_var.so timeout 900 python scripts/sweep_flow.py --steps 20 --configs "$SWEEP" 2>> $OUT/sweep.err | tee $OUT/sweep.jsonl | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print('W=%-3d G=%-8s %.1f steps/s  biteq=%s' % (d['W'], d['G'], d['steps_per_s'], d['bit_equal_to_first']))"
fi
if [ -n "$PROF" ]; then
  for g in $PROF; do
    TCFD_FLOW_PROF=1 TCFD_LIB=$PWD/torch-cfd_b200/libtcfd_var.so timeout 300 python scripts/sweep_flow.py --steps 10 --configs "$g" 2>> $OUT/sweep.err | tee -a $OUT/prof.jsonl
  done
fi
if [ -n "$NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ns2d_ -s 8 -c 4 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns2d_flow -s 8 -c 1 -o $OUT/prof_full -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
  ls -la $OUT
fi
tail -5 $OUT/bench.err
