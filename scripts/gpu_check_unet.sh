#!/bin/bash
# session 2: U-Net context cache: parity tests + denoise bench
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest unet + drivers" | tee -a $S
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_drivers_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 gpurun_out/pytest_unet.log
echo "== bench (denoise)" | tee -a $S
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e|denoise" gpurun_out/bench.err | tee -a $S
