mkdir -p gpurun_out; rm -f gpurun_out/r11_*
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/r11_test.log
for cfg in "256 18 750 64 64" "256 9 375 128 128" "256 5 188 256 256" "256 3 94 512 512"; do timeout 120 python scripts/prof_conv.py wgrad_patch $cfg >> gpurun_out/r11_prof.txt 2>&1; timeout 120 python scripts/prof_conv.py wgrad $cfg >> gpurun_out/r11_prof.txt 2>&1;  done
timeout 120 python scripts/prof_conv.py patch 256 3 94 512 512 >> gpurun_out/r11_prof.txt 2>&1
cat gpurun_out/r11_test.log; cat gpurun_out/r11_prof.txt
