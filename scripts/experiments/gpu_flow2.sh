#!/bin/bash
OUT=gpurun_out/flow2
mkdir -p $OUT
for W in 10 64; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ns2d_flow -s 3 -c 1 -o $OUT/flow_W$W -f \
  python scripts/sweep_flow.py --n 512 --batch 64 --steps 1 --configs 1:$W > $OUT/ncu_W$W.log 2>&1; echo "ncu W=$W rc=$?"
done
ls -la $OUT
