# compute-sanitizer over one B=2 train + score step of each net.  Logs under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out
for arch in resnet ecapa; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_step.py $arch 2 \
    > gpurun_out/sanitize_memcheck_$arch.log 2>&1; echo "memcheck $arch rc=$?" | tee -a gpurun_out/sanitize_memcheck_$arch.log
done
for arch in resnet ecapa; do
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_step.py $arch 2 \
    > gpurun_out/sanitize_racecheck_$arch.log 2>&1; echo "racecheck $arch rc=$?" | tee -a gpurun_out/sanitize_racecheck_$arch.log
done
for f in gpurun_out/sanitize_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_step|rc=" $f | tail -4; done
