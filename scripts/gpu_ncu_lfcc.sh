# one --set full capture of the tensor-core LFCC kernel + its summary (outputs under gpurun_out/)
mkdir -p gpurun_out
timeout 300 python bench.py --workload lfcc --steps 50 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lfcc ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
AIR_LFCC_IMPL=tc timeout 300 ncu --set full --clock-control none --import-source on -k regex:lfcc_tc_kernel -s 4 -c 1 -f -o gpurun_out/ncu_lfcc_tc python bench.py --workload lfcc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lfcc.log 2>&1
python scripts/ncu_src_summary.py gpurun_out/ncu_lfcc_tc.ncu-rep > gpurun_out/ncu_lfcc_tc_summary.txt 2>&1
cat gpurun_out/ncu_lfcc_tc_summary.txt
