mkdir -p gpurun_out; rm -f gpurun_out/r31_*
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_cli_gpu.py tests/test_kernels_gpu.py tests/test_ecapa_gpu.py -q 2>&1 | grep -E "passed|failed|FAILED|Error" | head > gpurun_out/r31_test.log
timeout 600 python bench.py --workload resnet_train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r31_bench_resnet.json 2> gpurun_out/r31_bench_resnet.err
AIR_OVERLAP_WGRAD=0 timeout 600 python bench.py --workload resnet_train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r31_bench_resnet_noov.json 2> gpurun_out/r31_bench_resnet_noov.err
cat gpurun_out/r31_test.log; python - <<'PY'
import json
for f in ("resnet","resnet_noov"):
    d=json.loads(open("gpurun_out/r31_bench_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"])
PY
