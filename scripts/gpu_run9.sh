mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -k patch > gpurun_out/r9_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r9_test.log
for cfg in "256 18 750 64 64" "256 9 375 128 128" "256 5 188 256 256"; do timeout 120 python scripts/prof_conv.py patch $cfg >> gpurun_out/r9_prof_conv.txt 2>&1; done
grep -v "^$" gpurun_out/r9_test.log | tail -30; cat gpurun_out/r9_prof_conv.txt
