# One gpurun call: the GPU parity suite, smoke, and the bench workloads (outputs under gpurun_out/).
# Also runs the still-unvalidated adversarial-head checks directly, so that their real output is kept (the pytest
# wrapper only reports xfail / xpass).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_full.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python tests/adv_gpu_checks.py > gpurun_out/adv_checks.log 2>&1; echo "adv rc=$?" >> gpurun_out/adv_checks.log
timeout 120 python bench.py --workload det > gpurun_out/bench_det.json 2> gpurun_out/bench_det.err
timeout 300 python scripts/bench_ingest.py > gpurun_out/ingest.json 2> gpurun_out/ingest.err
for w in resnet_train ecapa_train ecapa_score lfcc; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 300 python scripts/prof_step.py 256 > gpurun_out/percall_resnet.txt 2>&1
timeout 300 python scripts/prof_step.py 256 ecapa > gpurun_out/percall_ecapa.txt 2>&1
grep -v "^$" gpurun_out/test.log | tail -3; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/adv_checks.log
python - <<'PY'
import json
for f in ("resnet_train", "ecapa_train", "ecapa_score", "lfcc"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.3f ms" % d["ms_per_step"], "%.0f utt/s" % d["value"], "e2e %.0f" % d["e2e"]["value"],
              "roofline %.3f" % d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
