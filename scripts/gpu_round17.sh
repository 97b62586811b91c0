#!/bin/bash
mkdir -p gpurun_out
echo "== gemm check"; timeout 300 ./scripts/gemm_check.bin | cut -c1-150
echo "== probe (pair)"
timeout 120 ./scripts/gemm_probe.bin > gpurun_out/probe_pair.txt 2>&1; echo "rc=$?"
cat gpurun_out/probe_pair.txt | cut -c1-330
echo "== pytest unet"
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "rc=$?"
grep -E "rel err|relative error|passed|failed|Error|error|timeout" gpurun_out/pytest_unet.log | head -20
echo "== bench"
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
grep -E "profiled|timed|denoise" gpurun_out/bench.err
