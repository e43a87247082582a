mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
python -m pytest tests -m gpu -x -q -s > gpurun_out/r1_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_test.log
python bench.py --workload lfcc --steps 50 --warmup 5 > gpurun_out/r1_bench_lfcc.json 2> gpurun_out/r1_bench_lfcc.err
for f in 12 28 44 60 92 124; do python bench.py --workload lfcc --fseg $f --steps 50 --warmup 5 --no-cpu-baseline >> gpurun_out/r1_fseg_sweep.jsonl 2>> gpurun_out/r1_bench_lfcc.err; done
ncu --set full --clock-control none --import-source on -k regex:lfcc_kernel -s 6 -c 2 -o gpurun_out/r1_lfcc_prof python bench.py --workload lfcc --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu.log 2>&1
python __graft_entry__.py --smoke > gpurun_out/r1_smoke.log 2>&1
tail -5 gpurun_out/r1_test.log; cat gpurun_out/r1_bench_lfcc.json
