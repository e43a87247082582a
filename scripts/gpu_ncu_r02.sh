# Round-2 ncu evidence (one GPU): the launch list of a short default bench run (per-kernel share of the step) and
# --set full captures of the dominant kernels with their headline metrics as JSON.  ~8 GPU-minutes.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_r02.sh'
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --workload resnet_train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launch_summary.csv 2>/dev/null
cap() {  # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/r02_ncu_$name "$@" > gpurun_out/r02_ncu_$name.log 2>&1
}
cap patch_l1_stats conv_patch_kernel 3 python scripts/prof_conv.py patch_stats 256 18 750 64 64
cap patch_l1 conv_patch_kernel 3 python scripts/prof_conv.py patch 256 18 750 64 64
cap patch_l2 conv_patch_kernel 3 python scripts/prof_conv.py patch 256 9 375 128 128
cap patch_l3 conv_patch_kernel 3 python scripts/prof_conv.py patch 256 5 188 256 256
cap wgrad_patch_l1 conv3x3_wgrad_patch_kernel 3 python scripts/prof_conv.py wgrad_patch 256 18 750 64 64
cap wgrad_patch_l3 conv3x3_wgrad_patch_kernel 3 python scripts/prof_conv.py wgrad_patch 256 5 188 256 256
cap wgrad_patch_l4 conv3x3_wgrad_patch_kernel 3 python scripts/prof_conv.py wgrad_patch 256 3 94 512 512
cap s2_wgrad_l2 conv3x3_wgrad_patch_kernel 12 python scripts/prof_conv.py s2_wgrad 256 18 750 64 128
cap gemm_l4 conv_gemm_kernel 3 python scripts/prof_conv.py gemm 256 3 94 512 512
cap gemm_1x1_512 conv_gemm_kernel 3 python scripts/prof_conv.py gemm1x1 256 1 750 512 512
cap gemm_s2_l2 conv_gemm_kernel 3 python scripts/prof_conv.py gemm_s2 256 18 750 64 128
cap bn_bwd_apply_l1 bn_bwd_apply_kernel 3 python scripts/prof_conv.py bn_bwd 256 18 750 64 64
AIR_LFCC_IMPL=tc cap lfcc_tc lfcc_tc_kernel 4 python bench.py --workload lfcc --steps 3 --warmup 3 --no-cpu-baseline
args=""
for n in patch_l1_stats patch_l1 patch_l2 patch_l3 wgrad_patch_l1 wgrad_patch_l3 wgrad_patch_l4 s2_wgrad_l2 gemm_l4 gemm_1x1_512 gemm_s2_l2 bn_bwd_apply_l1 lfcc_tc; do
  [ -f gpurun_out/r02_ncu_$n.ncu-rep ] && args="$args $n=gpurun_out/r02_ncu_$n.ncu-rep"
done
python scripts/ncu_metrics.py $args > gpurun_out/r02_ncu_metrics.json
python scripts/ncu_src_summary.py gpurun_out/r02_ncu_lfcc_tc.ncu-rep > gpurun_out/r02_ncu_lfcc_tc_summary.txt 2>&1
python scripts/ncu_src_summary.py gpurun_out/r02_ncu_patch_l1_stats.ncu-rep > gpurun_out/r02_ncu_patch_l1_stats_summary.txt 2>&1
python scripts/ncu_src_summary.py gpurun_out/r02_ncu_patch_l1.ncu-rep > gpurun_out/r02_ncu_patch_l1_summary.txt 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r02_ncu_metrics.json'))
for k,v in d.items():
    print(k, {a: (round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ('time_us','tensor_pipe_active_pct','l1_lsu_wavefronts_pct','l1_tensor_operand_wavefronts_pct','l1_data_pipe_busy_pct','dram_throughput_pct','issue_active_pct','registers')})
"
head -24 gpurun_out/r02_launch_summary.csv
