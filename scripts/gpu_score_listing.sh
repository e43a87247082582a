#!/bin/bash
# Per-call listing of one ECAPA scoring step at the benchmark batch (B = 1024) + two repeats of the scoring bench.
mkdir -p gpurun_out
timeout 300 python scripts/prof_step.py 1024 ecapa score > gpurun_out/percall_ecapa_score.txt 2>&1
tail -3 gpurun_out/percall_ecapa_score.txt
for i in 1 2; do
  timeout 300 python bench.py --workload ecapa_score --steps 20 --warmup 5 > gpurun_out/bench_ecapa_score_rep$i.json 2> gpurun_out/bench_ecapa_score_rep$i.err
  cat gpurun_out/bench_ecapa_score_rep$i.json
done
