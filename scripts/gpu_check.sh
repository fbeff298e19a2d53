#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and full captures of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > $OUT/cpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 300 python bench.py --n 256 --steps 100 --no-cpu-baseline > $OUT/bench_256.json 2>> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
# full capture of the two substage kernels (skip warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns2d_ -s 12 -c 2 -o $OUT/prof_full -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
if [ -x scripts/microbench/fp32_pipe ]; then timeout 120 scripts/microbench/fp32_pipe > $OUT/microbench_fp32_pipe.txt 2>&1; fi
if [ -n "$SANITIZE" ]; then
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck.log 2>&1
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck.log 2>&1
  tail -3 $OUT/racecheck.log $OUT/memcheck.log
fi
ls -la $OUT
