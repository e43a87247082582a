# ncu evidence (one GPU): the launch list of a short default bench run and --set full captures of the dominant kernels.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_ncu.sh'
# Cost note (measured in round 1): ncu adds ~130 ms per launch even for the single-metric pass, so the launch list is
# capped with -c (3 warm-up steps + 2 timed steps of the ResNet workload are ~850 launches of ~170 per step); the
# three --set full captures take ~40 s each.  Budget ~6 GPU-minutes for the whole script.
mkdir -p gpurun_out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.csv 2>/dev/null
timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad_patch_kernel -s 3 -c 1 \
  -o gpurun_out/ncu_wgrad_patch_l1 python scripts/prof_conv.py wgrad_patch 256 18 750 64 64 > gpurun_out/ncu1.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv_patch_kernel -s 3 -c 1 \
  -o gpurun_out/ncu_patch_l1 python scripts/prof_conv.py patch 256 18 750 64 64 > gpurun_out/ncu2.log 2>&1
AIR_LFCC_IMPL=tc timeout 120 ncu --set full --clock-control none --import-source on -k regex:lfcc_tc_kernel -s 4 -c 1 \
  -o gpurun_out/ncu_lfcc_tc python bench.py --workload lfcc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out | grep -E "ncu|launch"
