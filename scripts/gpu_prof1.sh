mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches.csv python scripts/prof_step.py 256 > gpurun_out/r4_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 86 -c 16 -o gpurun_out/r4_conv_fprop python scripts/prof_step.py 256 > gpurun_out/r4_ncu_fprop.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 42 -c 21 -o gpurun_out/r4_conv_wgrad python scripts/prof_step.py 256 > gpurun_out/r4_ncu_wgrad.log 2>&1
ls -la gpurun_out | tail -8
