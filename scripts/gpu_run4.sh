mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r6_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r6_test.log
timeout 600 python bench.py --workload ecapa_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_ecapa.json 2> gpurun_out/r6_bench_ecapa.err
timeout 600 python bench.py --workload ecapa_score --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_ecapa_score.json 2> gpurun_out/r6_bench_ecapa_score.err
grep -v "^$" gpurun_out/r6_test.log | tail -40; cat gpurun_out/r6_bench_ecapa.json; tail -3 gpurun_out/r6_bench_ecapa.err;  cat gpurun_out/r6_bench_ecapa_score.json; tail -3 gpurun_out/r6_bench_ecapa_score.err
