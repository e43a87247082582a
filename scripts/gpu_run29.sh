mkdir -p gpurun_out; rm -f gpurun_out/r29_*
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r29_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r29_test.log
timeout 600 python bench.py --workload ecapa_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_ecapa.json 2> gpurun_out/r29_bench_ecapa.err
timeout 600 python bench.py --workload ecapa_score --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_ecapa_score.json 2> gpurun_out/r29_bench_ecapa_score.err
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_resnet.json 2> gpurun_out/r29_bench_resnet.err
timeout 300 python scripts/prof_step.py 256 ecapa > gpurun_out/r29_percall_ecapa.txt 2>&1
grep -v "^$" gpurun_out/r29_test.log | tail -3; python - <<'PY'
import json
for f in ("ecapa","ecapa_score","resnet"):
    d=json.loads(open("gpurun_out/r29_bench_%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
