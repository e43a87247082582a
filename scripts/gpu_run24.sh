mkdir -p gpurun_out; rm -f gpurun_out/r24_*
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/r24_test.log
timeout 600 python bench.py --workload resnet_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r24_bench_resnet.json 2> gpurun_out/r24_bench_resnet.err
timeout 300 python scripts/prof_step.py 256 > gpurun_out/r24_percall_resnet.txt 2>&1
# ncu: every launch of one default bench run with its device time (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r24_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r24_launches.log 2>&1
# ncu --set full of the three dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad_patch_kernel -s 3 -c 1 -o gpurun_out/r24_wgrad_patch_l1 python scripts/prof_conv.py wgrad_patch 256 18 750 64 64 > gpurun_out/r24_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_patch_kernel -s 3 -c 1 -o gpurun_out/r24_patch_l1 python scripts/prof_conv.py patch 256 18 750 64 64 > gpurun_out/r24_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lfcc_tc_kernel -s 4 -c 1 -o gpurun_out/r24_lfcc_tc python bench.py --workload lfcc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r24_ncu3.log 2>&1
cat gpurun_out/r24_test.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r24_bench_resnet.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
ls -la gpurun_out | grep r24
