mkdir -p gpurun_out; rm -f gpurun_out/r30_*
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_ecapa_kernels_gpu.py tests/test_ecapa_gpu.py tests/test_resnet_gpu.py -q 2>&1 | grep -E "passed|failed|FAILED|Error|assert" | head -30
