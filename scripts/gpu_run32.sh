mkdir -p gpurun_out; rm -f gpurun_out/r32_*
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r32_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r32_test.log
for w in resnet_train ecapa_train ecapa_score; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r32_bench_$w.json 2> gpurun_out/r32_bench_$w.err; done
timeout 300 python bench.py --workload lfcc --steps 50 --warmup 5 > gpurun_out/r32_bench_lfcc.json 2> gpurun_out/r32_bench_lfcc.err
grep -v "^$" gpurun_out/r32_test.log | tail -3; python - <<'PY'
import json
for f in ("resnet_train","ecapa_train","ecapa_score","lfcc"):
    try:
        d=json.loads(open("gpurun_out/r32_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f,"ERR",e)
PY
