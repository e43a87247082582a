mkdir -p gpurun_out; rm -f gpurun_out/r19_*
timeout 300 python -m pytest tests/test_lfcc_gpu.py -q -x -k "tc" 2>&1 | tail -4 > gpurun_out/r19.txt
for d in 0 30; do echo -n "dbg=$d " >> gpurun_out/r19.txt; AIR_LFCC_DBG=$d timeout 200 python bench.py --workload lfcc --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | sed -E 's/.*"ms_per_step": ([0-9.]+).*/\1/' >> gpurun_out/r19.txt; done
cat gpurun_out/r19.txt
