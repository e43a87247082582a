# fp32 parity mode + full-size parity on the GPU (outputs under gpurun_out/).
mkdir -p gpurun_out; rm -f gpurun_out/parity_full.json
timeout 1500 python -m pytest tests/test_parity_fp32_gpu.py -q -s > gpurun_out/parity_fp32.log 2>&1; echo "rc=$?" >> gpurun_out/parity_fp32.log
grep -E "passed|failed|error|fp32 mode|worst|losses|Error|assert" gpurun_out/parity_fp32.log | tail -30
if [ "$1" = "full" ]; then
timeout 1500 python -m pytest tests/test_parity_full_gpu.py -q -s > gpurun_out/parity_full.log 2>&1; echo "rc=$?" >> gpurun_out/parity_full.log
grep -E "passed|failed|error|B=256|B=1024|Error|assert" gpurun_out/parity_full.log | tail -30
fi
