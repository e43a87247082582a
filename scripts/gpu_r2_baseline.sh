mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test.log
tail -5 gpurun_out/test.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_resnet_train.json 2> gpurun_out/bench_resnet_train.err; tail -c 1500 gpurun_out/bench_resnet_train.json
bash scripts/gpu_sanitize.sh
