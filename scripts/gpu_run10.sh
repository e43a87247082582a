mkdir -p gpurun_out; rm -f gpurun_out/r10_dbg.txt
for d in 0 2 7 23; do echo "dbg=$d" >> gpurun_out/r10_dbg.txt; for cfg in "256 18 750 64 64" "256 9 375 128 128"; do AIR_PATCH_DBG=$d timeout 120 python scripts/prof_conv.py patch $cfg >> gpurun_out/r10_dbg.txt 2>&1; done; done
cat gpurun_out/r10_dbg.txt
