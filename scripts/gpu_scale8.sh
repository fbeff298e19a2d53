#!/bin/bash
# 8-GPU scaling check of the headline bench (weak scaling: 64 samples per GPU).
OUT=gpurun_out/r05_scale
mkdir -p $OUT
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_gpus$N.json 2> $OUT/bench_$N.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench_gpus$N.json')); print('n_gpus',d['n_gpus'],'steps/s=%.1f'%d['value'],'e2e=%.1f'%d['e2e']['value'], d['clocks'])"
tail -2 $OUT/bench_$N.err
