export TCFD_CHUNK_MB=100000
mkdir -p gpurun_out/r02f
for dbg in 0 1 2 3; do
  TCFD_DBG=$dbg TCFD_LIB=$PWD/torch-cfd_b200/libtcfd_b.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02f/bench_dbg$dbg.json 2>> gpurun_out/r02f/bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02f/bench_dbg$dbg.json"))
print("dbg $dbg", "steps/s=%.1f"%d["value"], {k:(round(v["us_per_launch"],1) if isinstance(v,dict) and v["us_per_launch"] else None) for k,v in d["kernels"].items() if isinstance(v,dict)})
PY
done
