mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test.log
timeout 300 python scripts/diag_resnet.py 4 > gpurun_out/r2_diag_resnet.log 2>&1
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 > gpurun_out/r2_bench_resnet.json 2> gpurun_out/r2_bench_resnet.err
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1
tail -30 gpurun_out/r2_test.log; cat gpurun_out/r2_bench_resnet.json; tail -5 gpurun_out/r2_bench_resnet.err; tail -3 gpurun_out/r2_smoke.log
