#!/bin/bash
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("steps/s=%.1f"%d["value"], "e2e=%.1f"%d["e2e"]["value"], "frac_step=%.3f"%d["roofline_step"]["frac"])
PY
for c in 2 8; do TCFD_HOST_CHUNKS=$c timeout 300 python bench.py --steps 20 --no-cpu-baseline 2>>$OUT/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('host chunks $c: e2e=%.1f'%d['e2e']['value'])"; done
timeout 600 python scripts/bench_fno3d.py > $OUT/bench_fno3d.json 2>> $OUT/bench.err; cat $OUT/bench_fno3d.json
tail -3 $OUT/bench.err
