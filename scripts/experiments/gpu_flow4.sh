#!/bin/bash
OUT=gpurun_out/flow4
mkdir -p $OUT
for G in 1,1,0 1,1,200 3,4,0 3,4,200 2,2,200 3,2,200 5,4,200; do
  echo "== G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 20 --configs 0:0,1:8,1:16,1:64 2> $OUT/sweep_512_$G.err | tee $OUT/sweep_512_$G.jsonl | cut -c1-200
  tail -2 $OUT/sweep_512_$G.err
done
for G in 1,1,0 1,1,200 3,2,200; do
  echo "== 256 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 256 --batch 64 --steps 50 --configs 0:0,1:16,1:64 2> $OUT/sweep_256_$G.err | tee $OUT/sweep_256_$G.jsonl | cut -c1-200
done
G=1,1,0
TCFD_FLOW_G=$G timeout 600 ncu --set full --clock-control none --import-source on -k regex:ns2d_flow -s 3 -c 1 -o $OUT/flow_v2_W64 -f \
  python scripts/sweep_flow.py --n 512 --batch 64 --steps 1 --configs 1:64 > $OUT/ncu_W64.log 2>&1; echo "ncu rc=$?"
