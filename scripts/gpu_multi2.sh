#!/bin/bash
# Multi-GPU confirmation visit (gpurun --gpus 8) with the round-end binaries: the three bench workloads under torchrun,
# launched exactly as the driver launches them.  Usage: bash scripts/gpu_multi2.sh <tag>
TAG=${1:-m2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) "${@:2}"; }
show() { python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['metric'], d['n_gpus'], 'GPUs: value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e') and round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3))"; }
timeout 300 python bench.py --steps 50 --no-cpu-baseline 2> $OUT/ns_1.err | tee $OUT/bench_512x64_1gpu.json | show
for n in 2 4 $NG; do
  run $n bench.py --gpus $n --steps 50 --no-cpu-baseline 2> $OUT/ns_$n.err | tee $OUT/bench_512x64_${n}gpu.json | show
done
run $NG bench.py --gpus $NG --workload fno3d_c5 --steps 10 --no-cpu-baseline 2> $OUT/fno_$NG.err | tee $OUT/bench_fno3d_${NG}gpu.json | show
run $NG bench.py --gpus $NG --workload sconv_c4 --steps 10 --no-cpu-baseline 2> $OUT/sconv_$NG.err | tee $OUT/bench_sconv_${NG}gpu.json | show
run $NG scripts/bench_trajectory.py --gather-to none --physical 1 --subsample 4 2> $OUT/c3.err | tee $OUT/c3_shard_physical_sub4.json | cut -c1-500
for f in $OUT/*.err; do echo "== $f"; tail -2 $f; done | tail -24
