#!/bin/bash
OUT=gpurun_out/flow11
mkdir -p $OUT
for G in 3,4,4 3,4,32; do
  echo "== 512 G=$G"
  TCFD_FLOW_G=$G timeout 300 python scripts/sweep_flow.py --n 512 --batch 64 --steps 100 --configs 1:64 2> $OUT/s512_$G.err | tee $OUT/s512_$G.jsonl | cut -c1-110
done
echo "== memcheck (flow schedule tests, small)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ns2d_gpu.py -m gpu -x -q -k "flow and 256" > $OUT/memcheck_ns2d.log 2>&1; echo "memcheck ns2d rc=$?"; tail -4 $OUT/memcheck_ns2d.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sconv_gpu.py -m gpu -x -q > $OUT/memcheck_sconv.log 2>&1; echo "memcheck sconv rc=$?"; tail -4 $OUT/memcheck_sconv.log
echo "== synccheck"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_ns2d_gpu.py -m gpu -x -q -k "flow and 256-20-6" > $OUT/synccheck_ns2d.log 2>&1; echo "synccheck rc=$?"; tail -4 $OUT/synccheck_ns2d.log
