#!/bin/bash
OUT=gpurun_out/flow7
mkdir -p $OUT
export TCFD_FLOW_VERBOSE=1
for P in 1 0; do
for G in 1,1,4 2,2,4 3,4,4; do
  echo "== persist=$P G=$G"
  TCFD_FLOW_PERSIST=$P TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 100 --configs 1:6,1:8,1:9,1:10,1:12,1:64 2> $OUT/s512_${P}_$G.err | tee $OUT/s512_${P}_$G.jsonl | cut -c1-120
  grep "tcfd: flow" $OUT/s512_${P}_$G.err | head -3
done
done
echo "== no-FFT persist W=8"
TCFD_FLOW_PERSIST=1 TCFD_FLOW_G=1,1,-4 timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 100 --configs 1:8,1:64 2> $OUT/nofft.err | cut -c1-120
for P in 1 0; do
  echo "== 256 persist=$P"
  TCFD_FLOW_PERSIST=$P TCFD_FLOW_G=1,1,8 timeout 300 python scripts/sweep_flow.py --n 256 --batch 64 --steps 200 --configs 1:16,1:24,1:32,1:64 2> $OUT/s256_$P.err | tee $OUT/s256_$P.jsonl | cut -c1-120
  grep "tcfd: flow" $OUT/s256_$P.err | head -2
done
