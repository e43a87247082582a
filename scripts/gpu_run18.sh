mkdir -p gpurun_out; rm -f gpurun_out/r18_*
for d in 0 1 2 4 8 16 24 25 27 31; do echo -n "dbg=$d " >> gpurun_out/r18.txt; AIR_LFCC_DBG=$d timeout 200 python bench.py --workload lfcc --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | sed -E 's/.*"ms_per_step": ([0-9.]+).*/\1/' >> gpurun_out/r18.txt; done
cat gpurun_out/r18.txt
