# LFCC iteration: parity tests of both implementations, the kernel-alone bench, the in-step figure
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lfcc_gpu.py -m gpu -q -x > gpurun_out/it_lfcc_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/it_lfcc_test.log
tail -4 gpurun_out/it_lfcc_test.log
timeout 300 python bench.py --workload lfcc --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/it_bench_lfcc.json 2> gpurun_out/it_bench_lfcc.err
python -c "
import json
d=json.loads(open('gpurun_out/it_bench_lfcc.json').read().strip().splitlines()[-1])
print('lfcc', d['ms_per_step'], 'ms', 'frac', d['roofline']['frac'], d['roofline'].get('fp32_exact_kernel'))
"
