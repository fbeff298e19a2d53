OUT=gpurun_out/s7; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sconv_planes -s 10 -c 2 -o $OUT/prof_planes -f python bench.py --workload sconv_c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu.log 2>&1
ls -la $OUT
