mkdir -p gpurun_out; rm -f gpurun_out/r21_*
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -q -x 2>&1 | tail -6 > gpurun_out/r21_test.log
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_resnet.json 2> gpurun_out/r21_bench_resnet.err
timeout 300 python scripts/prof_step.py 256 > gpurun_out/r21_percall_resnet.txt 2>&1
cat gpurun_out/r21_test.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r21_bench_resnet.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
tail -3 gpurun_out/r21_bench_resnet.err
