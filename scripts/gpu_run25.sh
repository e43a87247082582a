mkdir -p gpurun_out; rm -f gpurun_out/r25_*
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/r25_test.log
for cfg in "256 18 750 64 64" "256 9 375 128 128"; do for k in patch wgrad_patch; do timeout 120 python scripts/prof_conv.py $k $cfg >> gpurun_out/r25_prof.txt 2>&1; done; done
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r25_bench_resnet.json 2> gpurun_out/r25_bench_resnet.err
cat gpurun_out/r25_test.log gpurun_out/r25_prof.txt; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r25_bench_resnet.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
