# 8-GPU (or N-GPU) bench: the default exchange / compute SM split and the round-1 setting, plus the DDP gradient check
N=${1:-8}
mkdir -p gpurun_out
run() {
  local tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
    bench.py --gpus $N --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/scale_${N}gpu_$tag.json 2> gpurun_out/scale_${N}gpu_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/scale_${N}gpu_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'N=%d' % d['n_gpus'], '%.3f ms' % d['ms_per_step'], '%.0f utt/s' % d['value'], 'e2e %.0f' % d['e2e']['value'], d['clocks'])
" || tail -5 gpurun_out/scale_${N}gpu_$tag.err
}
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/scale_1gpu.json 2> gpurun_out/scale_1gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/scale_1gpu.json').read().strip().splitlines()[-1])
print('1gpu', '%.3f ms' % d['ms_per_step'], '%.0f utt/s' % d['value'], d['clocks'])
"
run default
run old8 NCCL_MAX_CTAS=8 AIR_RESERVE_SMS=0
