mkdir -p gpurun_out/r04_2gpu
N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/multi_gpu_check.py 2>&1 | grep -E "world=|MULTI_GPU|Error|error" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r04_2gpu/bench_gpus$N.json 2> gpurun_out/r04_2gpu/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r04_2gpu/bench_gpus$N.json')); print('n_gpus',d['n_gpus'],'steps/s=%.1f'%d['value'],'e2e=%.1f'%d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 scripts/bench_fno3d.py 2>>gpurun_out/r04_2gpu/bench.err | tee gpurun_out/r04_2gpu/bench_fno3d_gpus$N.json
tail -3 gpurun_out/r04_2gpu/bench.err
