mkdir -p gpurun_out; rm -f gpurun_out/r23_*
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r23_gpus.txt 2>&1
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r23_bench_8gpu.json 2> gpurun_out/r23_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r23_bench_4gpu.json 2> gpurun_out/r23_bench_4gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r23_bench_8gpu.json","gpurun_out/r23_bench_4gpu.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r23_bench_8gpu.err
