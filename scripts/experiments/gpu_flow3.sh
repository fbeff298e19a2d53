#!/bin/bash
OUT=gpurun_out/flow3
mkdir -p $OUT
timeout 600 python -m pytest tests/test_ns2d_gpu.py -m gpu -x -q > $OUT/pytest_ns2d.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_ns2d.log
for G in 1,1 3,4 3,2 5,4 5,8 9,8; do
  echo "== G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 20 --multi 10 --configs 0:0,1:8,1:10,1:16,1:64 2> $OUT/sweep_512_$G.err | tee $OUT/sweep_512_$G.jsonl | cut -c1-200
  tail -2 $OUT/sweep_512_$G.err
done
for G in 3,2 3,4 5,4; do
  echo "== 256 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 256 --batch 64 --steps 50 --configs 0:0,1:32,1:64 2> $OUT/sweep_256_$G.err | tee $OUT/sweep_256_$G.jsonl | cut -c1-200
done
