#!/bin/bash
# First GPU visit of the dataflow schedule: ns2d gpu tests (flow on by default), then the schedule sweep.
OUT=gpurun_out/flow1
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_ns2d_gpu.py -m gpu -x -q > $OUT/pytest_ns2d.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_ns2d.log
tail -5 $OUT/pytest_ns2d.log
timeout 600 python scripts/sweep_flow.py --n 512 --batch 64 --steps 20 --multi 10 > $OUT/sweep_512.jsonl 2> $OUT/sweep_512.err; echo "sweep512 rc=$?"
cat $OUT/sweep_512.jsonl; tail -3 $OUT/sweep_512.err
timeout 300 python scripts/sweep_flow.py --n 256 --batch 64 --steps 50 --configs 0:0,1:8,1:16,1:32,1:64 > $OUT/sweep_256.jsonl 2> $OUT/sweep_256.err; echo "sweep256 rc=$?"
cat $OUT/sweep_256.jsonl; tail -3 $OUT/sweep_256.err
