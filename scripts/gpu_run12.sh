mkdir -p gpurun_out; rm -f gpurun_out/r12_*
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r12_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r12_test.log
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r12_bench_resnet.json 2> gpurun_out/r12_bench_resnet.err
timeout 300 python scripts/prof_step.py 256 > gpurun_out/r12_percall_resnet.txt 2>&1
grep -v "^$" gpurun_out/r12_test.log | tail -30; cat gpurun_out/r12_bench_resnet.json; tail -3 gpurun_out/r12_bench_resnet.err
