mkdir -p gpurun_out; rm -f gpurun_out/r28_*
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_ecapa_gpu.py tests/test_ecapa_kernels_gpu.py tests/test_cli_gpu.py -q -x > gpurun_out/r28_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r28_test.log
timeout 600 python bench.py --workload ecapa_score --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r28_bench_ecapa_score.json 2> gpurun_out/r28_bench_ecapa_score.err
timeout 600 python bench.py --workload ecapa_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r28_bench_ecapa.json 2> gpurun_out/r28_bench_ecapa.err
grep -v "^$" gpurun_out/r28_test.log | tail -3; python - <<'PY'
import json
for f in ("ecapa_score","ecapa"):
    d=json.loads(open("gpurun_out/r28_bench_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
