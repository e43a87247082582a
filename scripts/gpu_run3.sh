mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ecapa_kernels_gpu.py tests/test_ecapa_gpu.py tests/test_resnet_gpu.py -q -s > gpurun_out/r5_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r5_test.log
timeout 600 python bench.py --workload ecapa_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench_ecapa.json 2> gpurun_out/r5_bench_ecapa.err
timeout 300 python scripts/prof_step.py 256 ecapa > gpurun_out/r5_percall_ecapa.txt 2>&1
grep -v "^$" gpurun_out/r5_test.log | tail -40; cat gpurun_out/r5_bench_ecapa.json; tail -5 gpurun_out/r5_bench_ecapa.err
