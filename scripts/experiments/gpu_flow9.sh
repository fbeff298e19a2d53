#!/bin/bash
OUT=gpurun_out/flow9
mkdir -p $OUT
for G in 3,4,4 3,4,-4 3,4,-5 3,4,-9 3,4,-17 3,4,-29; do
  echo "== 512 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 100 --configs 1:64 2> $OUT/s512_$G.err | tee $OUT/s512_$G.jsonl | cut -c1-110
  tail -1 $OUT/s512_$G.err
done
