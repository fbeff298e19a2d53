mkdir -p gpurun_out/r03a
timeout 900 python -m pytest tests/test_sconv_gpu.py -m gpu -x -q > gpurun_out/r03a/pytest_sconv.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r03a/pytest_sconv.log
timeout 600 python scripts/bench_sconv.py > gpurun_out/r03a/bench_sconv.json 2> gpurun_out/r03a/bench_sconv.err; echo "bench rc=$?"; cat gpurun_out/r03a/bench_sconv.json; tail -5 gpurun_out/r03a/bench_sconv.err
