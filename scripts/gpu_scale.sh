# Weak-scaling bench of the default workload on N GPUs of one box (N = 2, 4 or 8) plus the DDP gradient check:
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale.sh 8'
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 \
  scripts/ddp_check.py resnet > gpurun_out/ddp_check_resnet.log 2>&1
cat gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/ddp_check_resnet.log
