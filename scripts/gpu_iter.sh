# One iteration on the GPU: the conv / resnet parity tests, the ResNet bench with the fusion on and off, a per-call listing.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_iter.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/it_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/it_test.log
tail -5 gpurun_out/it_test.log
timeout 300 python bench.py --workload resnet_train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/it_bench_fused.json 2> gpurun_out/it_bench_fused.err
AIR_FUSE_BN_BWD=0 timeout 300 python bench.py --workload resnet_train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/it_bench_unfused.json 2> gpurun_out/it_bench_unfused.err
timeout 300 python scripts/prof_step.py 256 > gpurun_out/it_percall_resnet.txt 2>&1
python - <<'PY'
import json
for f in ("fused", "unfused"):
    try:
        d = json.loads(open("gpurun_out/it_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.3f ms" % d["ms_per_step"], "%.0f utt/s" % d["value"], "roofline %.3f" % d["roofline"]["frac"],
              {k: v["ms_per_step"] for k, v in d["kernels"].items() if v["ms_per_step"] > 0.2})
    except Exception as e:
        print(f, "ERR", e)
PY
