#!/bin/bash
OUT=gpurun_out/flow8
mkdir -p $OUT
export TCFD_FLOW_PERSIST=0
for G in 3,4,4 3,4,2 1,1,4 1,1,2 3,4,-2; do
  echo "== 512 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 100 --configs 1:64,1:16 2> $OUT/s512_$G.err | tee $OUT/s512_$G.jsonl | cut -c1-120
  tail -1 $OUT/s512_$G.err
done
for G in 1,1,8 1,1,2 3,4,2; do
  echo "== 256 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 256 --batch 64 --steps 200 --configs 1:64 2> $OUT/s256_$G.err | tee $OUT/s256_$G.jsonl | cut -c1-120
done
