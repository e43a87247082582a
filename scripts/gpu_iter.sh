# One iteration on the GPU: kernel / net parity tests, the ResNet bench (plus a variant with an environment switch), a
# serialised per-call listing.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_iter.sh [VAR=value]'
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py tests/test_kernels_gpu.py tests/test_ecapa_kernels_gpu.py tests/test_ecapa_gpu.py -m gpu -q -x > gpurun_out/it_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/it_test.log
tail -4 gpurun_out/it_test.log
timeout 300 python bench.py --workload resnet_train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/it_bench_a.json 2> gpurun_out/it_bench_a.err
if [ -n "$1" ]; then
  env "$1" timeout 300 python bench.py --workload resnet_train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/it_bench_b.json 2> gpurun_out/it_bench_b.err
fi
AIR_OVERLAP_WGRAD=0 timeout 300 python scripts/prof_step.py 256 > gpurun_out/it_percall_resnet.txt 2>&1
python - <<'PY'
import json
for f in ("a", "b"):
    try:
        d = json.loads(open("gpurun_out/it_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.3f ms" % d["ms_per_step"], "%.0f utt/s" % d["value"], "roofline %.3f" % d["roofline"]["frac"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.2})
    except Exception as e:
        print(f, "ERR", e)
PY
AIR_OVERLAP_WGRAD=0 timeout 300 python scripts/prof_step.py 256 ecapa > gpurun_out/it_percall_ecapa.txt 2>&1
for w in ecapa_train ecapa_score; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/it_bench_$w.json 2> gpurun_out/it_bench_$w.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/it_bench_$w.json').read().strip().splitlines()[-1])
print('$w', '%.3f ms' % d['ms_per_step'], '%.0f utt/s' % d['value'], 'roofline %.3f' % d['roofline']['frac'], {k: v['ms_per_step'] for k, v in d.get('kernels', {}).items() if v['ms_per_step'] > 0.3})
"
done
