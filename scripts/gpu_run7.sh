mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r7_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r7_test.log
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_resnet.json 2> gpurun_out/r7_bench_resnet.err
timeout 300 python scripts/prof_step.py 256 > gpurun_out/r7_percall_resnet.txt 2>&1
for k in patch gemm; do for cfg in "256 18 750 64 64" "256 9 375 128 128" "256 5 188 256 256" "256 3 94 512 512"; do timeout 120 python scripts/prof_conv.py $k $cfg >> gpurun_out/r7_prof_conv.txt 2>&1; done; done
grep -v "^$" gpurun_out/r7_test.log | tail -30; cat gpurun_out/r7_bench_resnet.json; tail -3 gpurun_out/r7_bench_resnet.err; tail -30 gpurun_out/r7_prof_conv.txt
