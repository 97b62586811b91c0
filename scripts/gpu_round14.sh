#!/bin/bash
mkdir -p gpurun_out
echo "== unet tiny check (fused attention)"
timeout 300 python scripts/unet_tiny_check.py 16 > gpurun_out/unet16.log 2>&1; echo "rc=$?"
tail -16 gpurun_out/unet16.log
echo "== pytest unet"
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "rc=$?"
grep -E "rel err|relative error|passed|failed|Error|error|timeout" gpurun_out/pytest_unet.log | head -20
echo "== bench"
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
grep -E "profiled|timed|denoise" gpurun_out/bench.err
