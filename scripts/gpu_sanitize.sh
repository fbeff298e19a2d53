#!/bin/bash
# compute-sanitizer visit for the round-2 kernels: memcheck / racecheck / synccheck of the dataflow-schedule tests (blocked
# advection rows, rank-4 TMA loads) and of the spectral-convolution tests (pruned inverse, LDGSTS staging, shared-tile mix).
# The dependency time-out of the dataflow kernel is disabled (TCFD_FLOW_TIMEOUT_S=0): the tools slow the kernels 10-100 x.
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export TCFD_FLOW_TIMEOUT_S=0
K_NS="flow_schedule_bit_identical or vs_oracle_all_sizes"
K_SC="sizes or generations"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_ns2d_gpu.py -x -q -k "$K_NS" > $OUT/memcheck_ns2d.log 2>&1; tail -2 $OUT/memcheck_ns2d.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_ns2d_gpu.py -x -q -k "vs_oracle_all_sizes" > $OUT/racecheck_ns2d.log 2>&1; tail -2 $OUT/racecheck_ns2d.log
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_ns2d_gpu.py -x -q -k "vs_oracle_all_sizes" > $OUT/synccheck_ns2d.log 2>&1; tail -2 $OUT/synccheck_ns2d.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_sconv_gpu.py -x -q -k "$K_SC" > $OUT/memcheck_sconv.log 2>&1; tail -2 $OUT/memcheck_sconv.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_sconv_gpu.py -x -q -k "sizes" > $OUT/racecheck_sconv.log 2>&1; tail -2 $OUT/racecheck_sconv.log
grep -c "WARNING\|ERROR" $OUT/*.log
