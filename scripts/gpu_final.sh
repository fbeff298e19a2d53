#!/bin/bash
# Round-end validation visit: the full GPU test tier, the three bench lines (ours) + the reference arm of the headline,
# launch lists and one ncu --set full capture of the headline kernel.  Usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_512.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload sconv_c4 --steps 10 > $OUT/bench_sconv.json 2>> $OUT/bench.err
timeout 600 python bench.py --workload fno3d_c5 --steps 10 > $OUT/bench_fno3d.json 2>> $OUT/bench.err
timeout 300 python bench.py --n 256 --steps 100 --no-cpu-baseline --no-e2e > $OUT/bench_256.json 2>> $OUT/bench.err
python - <<PY
import json
for f in ("bench_512", "bench_reference", "bench_sconv", "bench_fno3d", "bench_256"):
    try:
        d = json.load(open("$OUT/%s.json" % f))
        print(f, "value=%.2f" % d["value"], d["unit"], "ms/step=%.3f" % d["ms_per_step"], "frac=", d.get("roofline", {}).get("frac"), "e2e=", d.get("e2e") and round(d["e2e"]["value"], 2), "cpu=", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ns2d_ -s 8 -c 4 --csv --log-file $OUT/launches_512x64.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ns2d_flow -s 8 -c 1 -o $OUT/prof_full -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 26 --csv --log-file $OUT/launches_fno3d.csv python bench.py --workload fno3d_c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sconv -s 25 -c 11 --csv --log-file $OUT/launches_sconv.csv python bench.py --workload sconv_c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu3.log 2>&1
ls -la $OUT; tail -3 $OUT/bench.err
