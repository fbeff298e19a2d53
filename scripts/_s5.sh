OUT=gpurun_out/s5; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "not horizon" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 300 python scripts/bench_c1.py > $OUT/bench_c1.json 2> $OUT/c1.err; cat $OUT/bench_c1.json; tail -2 $OUT/c1.err
timeout 600 python bench.py --workload sconv_c4 --steps 10 --no-cpu-baseline > $OUT/bench_sconv.json 2> $OUT/bench_sconv.err; python -c "
import json; d=json.load(open('$OUT/bench_sconv.json')); print('sconv fwd+bwd ms', d['ms_per_step'], 'fwd ms', d['config']['ms_forward_only'], 'frac', d['roofline']['frac'])"
timeout 600 python bench.py --workload fno3d_c5 --steps 10 --no-cpu-baseline > $OUT/bench_fno3d.json 2> $OUT/bench_fno3d.err; python -c "
import json; d=json.load(open('$OUT/bench_fno3d.json')); print('fno3d ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
