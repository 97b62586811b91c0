#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== unet tiny check 16" | tee -a $S
timeout 300 python scripts/unet_tiny_check.py 16 > gpurun_out/unet16.log 2>&1; echo "rc=$?" | tee -a $S
tail -25 gpurun_out/unet16.log
echo "== unet tiny check 32" | tee -a $S
timeout 300 python scripts/unet_tiny_check.py 32 > gpurun_out/unet32.log 2>&1; echo "rc=$?" | tee -a $S
tail -20 gpurun_out/unet32.log
if ! grep -q "eps  " gpurun_out/unet16.log; then
echo "== memcheck" | tee -a $S
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/unet_tiny_check.py 16 > gpurun_out/unet16_memcheck.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "Invalid|Error|at 0x|by thread|Saved host|unet_|uce::" gpurun_out/unet16_memcheck.log | head -40
fi
