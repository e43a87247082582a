N=${1:-2}
mkdir -p gpurun_out
AIR_DDP_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 \
  bench.py --gpus $N --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/ddp_trace_${N}gpu.json 2> gpurun_out/ddp_trace_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/ddp_trace_${N}gpu.json').read().strip().splitlines()[-1])
print('N=%d' % d['n_gpus'], '%.3f ms' % d['ms_per_step'], 'exposed exchange %.3f ms/step' % d.get('exchange_exposed_ms_per_step', -1))
" || tail -5 gpurun_out/ddp_trace_${N}gpu.err
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu %.3f ms' % d['ms_per_step'])"
