mkdir -p gpurun_out; rm -f gpurun_out/r13_*
timeout 900 python -m pytest tests/test_conv_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r13_test.log
timeout 900 python -m pytest tests/test_resnet_gpu.py -q -x 2>&1 | tail -25 >> gpurun_out/r13_test.log
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r13_bench_resnet.json 2> gpurun_out/r13_bench_resnet.err
timeout 300 python scripts/prof_step.py 256 > gpurun_out/r13_percall_resnet.txt 2>&1
cat gpurun_out/r13_test.log; cat gpurun_out/r13_bench_resnet.json; tail -3 gpurun_out/r13_bench_resnet.err
