mkdir -p gpurun_out; rm -f gpurun_out/r26_*
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r26_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r26_test.log
timeout 600 python bench.py --workload ecapa_score --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r26_bench_ecapa_score.json 2> gpurun_out/r26_bench_ecapa_score.err
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r26_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r26_smoke.log
grep -v "^$" gpurun_out/r26_test.log | tail -5; tail -4 gpurun_out/r26_smoke.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r26_bench_ecapa_score.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
