#!/bin/bash
# Full GPU-box visit: all gpu tests, smoke, default bench (+ reference arm), C2 bench, spectral-conv and
# FNO3d benches, ncu launch lists and one full capture of the dominant kernel.
TAG=${1:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > $OUT/cpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 300 python bench.py --n 256 --steps 100 --no-cpu-baseline > $OUT/bench_256.json 2>> $OUT/bench.err
cut -c1-400 $OUT/bench_256.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
tail -3 $OUT/bench.err
timeout 300 python scripts/bench_sconv.py 2>> $OUT/bench.err | tee $OUT/bench_sconv.jsonl | cut -c1-300
timeout 300 python scripts/bench_fno3d.py 2>> $OUT/bench.err | tee $OUT/bench_fno3d.json
# launch lists (cold-cache, serialised times: shares of the step, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sconv -s 60 -c 12 --csv --log-file $OUT/launches_sconv.csv \
  python scripts/bench_sconv.py --iters 4 > $OUT/sconv_under_ncu.log 2>&1; echo "ncu sconv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ns2d_flow -s 5 -c 1 -o $OUT/flow_full -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
