mkdir -p gpurun_out; rm -f gpurun_out/r17_*
timeout 300 python -m pytest tests/test_lfcc_gpu.py -q -x -k "tc" > gpurun_out/r17_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r17_test.log
timeout 200 python bench.py --workload lfcc --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r17_bench_lfcc.json 2> gpurun_out/r17_bench_lfcc.err
AIR_LFCC_IMPL=fft timeout 200 python bench.py --workload lfcc --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r17_bench_lfcc_fft.json 2>> gpurun_out/r17_bench_lfcc.err
grep -v "^$" gpurun_out/r17_test.log | tail -30; cat gpurun_out/r17_bench_lfcc.json; cat gpurun_out/r17_bench_lfcc_fft.json; tail -5 gpurun_out/r17_bench_lfcc.err
