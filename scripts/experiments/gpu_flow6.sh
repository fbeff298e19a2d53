#!/bin/bash
OUT=gpurun_out/flow6
mkdir -p $OUT
echo "== no-FFT timing experiment (3,4,-4) vs default"
for G in 3,4,4 3,4,-4; do
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 100 --configs 1:64 2> $OUT/nofft_$G.err | tee $OUT/nofft_$G.jsonl | cut -c1-160
done
echo "== e2e host chunks"
for c in 1 2 4 8 16; do
  TCFD_HOST_CHUNKS=$c timeout 300 python bench.py --steps 20 --no-cpu-baseline 2>> $OUT/e2e.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
# launch list (cold-cache, serialised times: shares of the step, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ns2d_flow -s 5 -c 1 -o $OUT/flow_full -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
