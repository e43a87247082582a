mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv3x3_patch_kernel -s 3 -c 1 -o gpurun_out/r8_patch_l1 python scripts/prof_conv.py patch 256 18 750 64 64 > gpurun_out/r8_ncu_patch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 3 -c 1 -o gpurun_out/r8_wgrad_l1 python scripts/prof_conv.py wgrad 256 18 750 64 64 > gpurun_out/r8_ncu_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 3 -c 1 -o gpurun_out/r8_gemm_l4 python scripts/prof_conv.py gemm 256 3 94 512 512 > gpurun_out/r8_ncu_gemm.log 2>&1
ls -la gpurun_out | tail; tail -3 gpurun_out/r8_ncu_patch.log
