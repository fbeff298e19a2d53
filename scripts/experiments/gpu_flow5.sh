#!/bin/bash
OUT=gpurun_out/flow5
mkdir -p $OUT
timeout 600 python -m pytest tests/test_ns2d_gpu.py -m gpu -x -q -k flow > $OUT/pytest_flow.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_flow.log
for G in 1,1,0 1,1,4 3,4,0 3,4,4 2,2,4; do
  echo "== G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 20 --configs 0:0,1:16,1:64 2> $OUT/sweep_512_$G.err | tee $OUT/sweep_512_$G.jsonl | cut -c1-200
  tail -2 $OUT/sweep_512_$G.err
done
for G in 1,1,0 1,1,4 2,2,4; do
  echo "== 256 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 256 --batch 64 --steps 50 --configs 0:0,1:64 2> $OUT/sweep_256_$G.err | tee $OUT/sweep_256_$G.jsonl | cut -c1-200
done
