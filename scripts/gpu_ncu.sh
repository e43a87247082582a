# ncu evidence (one GPU): the launch list of a short default bench run and --set full captures of the dominant kernels.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_ncu.sh'
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad_patch_kernel -s 3 -c 1 \
  -o gpurun_out/ncu_wgrad_patch_l1 python scripts/prof_conv.py wgrad_patch 256 18 750 64 64 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_patch_kernel -s 3 -c 1 \
  -o gpurun_out/ncu_patch_l1 python scripts/prof_conv.py patch 256 18 750 64 64 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lfcc_tc_kernel -s 4 -c 1 \
  -o gpurun_out/ncu_lfcc_tc python bench.py --workload lfcc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out | grep ncu
