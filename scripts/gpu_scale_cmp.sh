# N-GPU bench of the default workload, three configurations of the exchange / compute SM split, plus the 1-GPU run of the
# same box:   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_scale_cmp.sh 2'
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag env...
  local tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus $N --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/scale_${N}gpu_$tag.json 2> gpurun_out/scale_${N}gpu_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/scale_${N}gpu_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'N=%d' % d['n_gpus'], '%.3f ms' % d['ms_per_step'], '%.0f utt/s' % d['value'], 'e2e %.0f' % d['e2e']['value'], d['clocks'])
"
}
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/scale_1gpu.json 2> gpurun_out/scale_1gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/scale_1gpu.json').read().strip().splitlines()[-1])
print('1gpu', '%.3f ms' % d['ms_per_step'], '%.0f utt/s' % d['value'], d['clocks'])
"
run reserve4 NCCL_MAX_CTAS=4
run reserve2 NCCL_MAX_CTAS=2
run old8 NCCL_MAX_CTAS=8 AIR_RESERVE_SMS=0
run reserve8 NCCL_MAX_CTAS=8
