#!/bin/bash
# One GPU-box visit for hot path B: spectral-conv parity tests, C4 / C5 bench lines, A/B of the plane-kernel generations.
TAG=${1:-sc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_sconv_gpu.py tests/test_sfno_gpu.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -4 $OUT/pytest.log
run() {  # tag, env assignments...
  local tag=$1; shift
  env "$@" timeout 300 python bench.py --workload sconv_c4 --no-cpu-baseline --no-e2e > $OUT/sconv_$tag.json 2>> $OUT/bench.err
  env "$@" timeout 300 python bench.py --workload fno3d_c5 --no-cpu-baseline --no-e2e > $OUT/fno3d_$tag.json 2>> $OUT/bench.err
}
run new TCFD_X=0
run old TCFD_SCONV_PLANES=2 TCFD_SCONV_XAXIS=1 TCFD_SCONV_MIX=1
[ -n "$EXTRA" ] && run extra $EXTRA
python - <<PY
import json
for f in ["sconv_new","sconv_old","sconv_extra","fno3d_new","fno3d_old","fno3d_extra"]:
    try:
        d=json.load(open("$OUT/"+f+".json")); print(f, "ms/step=%.3f"%d["ms_per_step"], "fwd=", d["config"].get("ms_forward_only"), "frac=%.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
if [ -n "$NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sconv_ -s 39 -c 13 --csv --log-file $OUT/launches_sconv.csv \
    python bench.py --workload sconv_c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launch.log 2>&1
  if [ "$NCU" != "list" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$NCU" -s 4 -c 2 -o $OUT/prof -f \
    python bench.py --workload sconv_c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
  fi
fi
if [ -n "$SANITIZE" ]; then
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_sconv_gpu.py -x -q -k "sizes" > $OUT/memcheck.log 2>&1; tail -3 $OUT/memcheck.log
fi
tail -3 $OUT/bench.err
