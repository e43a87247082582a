mkdir -p gpurun_out; rm -f gpurun_out/r20_*
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r20_bench_2gpu.json 2> gpurun_out/r20_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 scripts/ddp_check.py resnet > gpurun_out/r20_ddp_check_resnet.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r20_bench_1gpu.json 2> gpurun_out/r20_bench_1gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r20_bench_1gpu.json","gpurun_out/r20_bench_2gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("roofline_lfcc"), d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
tail -4 gpurun_out/r20_ddp_check_resnet.log; tail -3 gpurun_out/r20_bench_2gpu.err
