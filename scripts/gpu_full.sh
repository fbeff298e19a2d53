#!/bin/bash
# Full GPU-box visit: all gpu tests, smoke, default bench (+ reference arm), C2 bench.
TAG=${1:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export TCFD_CHUNK_MB=${TCFD_CHUNK_MB:-100000}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > $OUT/cpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 300 python bench.py --n 256 --steps 100 --no-cpu-baseline > $OUT/bench_256.json 2>> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
tail -3 $OUT/bench.err
