#!/bin/bash
# One multi-GPU visit (gpurun --gpus 8): config C3 end to end with the device-side post-processing and per-rank
# collection, and the path-B workloads at 2 / 4 / 8 GPUs.  Usage: bash scripts/gpu_multi.sh <tag>
TAG=${1:-m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "${@:2}"; }
# C3: 2000 steps, 512^2, 32 samples per GPU, 100 snapshots; every rank post-processes and copies ITS shard
run $NG scripts/bench_trajectory.py --gather-to none --physical 1 --subsample 1 2> $OUT/c3a.err | tee $OUT/c3_shard_physical.json | cut -c1-600
run $NG scripts/bench_trajectory.py --gather-to none --physical 1 --subsample 4 2> $OUT/c3b.err | tee $OUT/c3_shard_physical_sub4.json | cut -c1-600
run $NG scripts/bench_trajectory.py --gather-to 0 --physical 1 --subsample 4 2> $OUT/c3c.err | tee $OUT/c3_gather0_physical_sub4.json | cut -c1-600
for n in 2 4 $NG; do
  run $n bench.py --gpus $n --workload fno3d_c5 --steps 10 --no-cpu-baseline 2> $OUT/fno_$n.err | tee $OUT/bench_fno3d_${n}gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fno3d_c5', d['n_gpus'], 'GPUs: steps/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))"
done
run $NG bench.py --gpus $NG --workload sconv_c4 --steps 10 --no-cpu-baseline 2> $OUT/sconv_$NG.err | tee $OUT/bench_sconv_${NG}gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sconv_c4', d['n_gpus'], 'GPUs: steps/s', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))"
for f in $OUT/*.err; do echo "== $f"; tail -2 $f; done | tail -30
