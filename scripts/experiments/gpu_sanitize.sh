#!/bin/bash
OUT=gpurun_out/sanitize
mkdir -p $OUT
for tool in racecheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_ns2d_gpu.py -m gpu -x -q -k "flow and 256-20-6" > $OUT/${tool}_ns2d.log 2>&1; echo "$tool ns2d rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/${tool}_ns2d.log | tail -3
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_sconv_gpu.py -m gpu -x -q -k "golden or fused" > $OUT/${tool}_sconv.log 2>&1; echo "$tool sconv rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/${tool}_sconv.log | tail -3
done
