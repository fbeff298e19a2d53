OUT=gpurun_out/s8; mkdir -p $OUT
timeout 900 python -m pytest tests/test_sconv_gpu.py tests/test_sfno_gpu.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --workload sconv_c4 --steps 10 --no-cpu-baseline > $OUT/bench_sconv.json 2> $OUT/bench_sconv.err; python -c "
import json; d=json.load(open('$OUT/bench_sconv.json')); print('sconv fwd+bwd ms', d['ms_per_step'], 'fwd ms', d['config']['ms_forward_only'], 'frac', d['roofline']['frac'])"
timeout 600 python bench.py --workload fno3d_c5 --steps 10 --no-cpu-baseline > $OUT/bench_fno3d.json 2> $OUT/bench_fno3d.err; python -c "
import json; d=json.load(open('$OUT/bench_fno3d.json')); print('fno3d ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sconv -s 25 -c 11 --csv --log-file $OUT/launches_sconv.csv python bench.py --workload sconv_c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/s8/launches_sconv.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cur = {}
for r in rows[hdr + 1:]:
    key = (r[0], r[4].split('(')[0][:50]); cur.setdefault(key, {})[r[12]] = float(r[14])
for (i, name), v in cur.items():
    print('%3s %-52s %8.1f us  dram %.2f GB' % (i, name, v.get('gpu__time_duration.sum', 0) / 1e3, (v.get('dram__bytes_read.sum', 0) + v.get('dram__bytes_write.sum', 0)) / 1e9))
PY
