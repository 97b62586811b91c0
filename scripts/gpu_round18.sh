#!/bin/bash
mkdir -p gpurun_out
echo "== gemm check"; timeout 300 ./scripts/gemm_check.bin | tail -3 | cut -c1-150
echo "== pytest unet"
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "rc=$?"
grep -E "rel err|relative error|passed|failed|Error|error|timeout" gpurun_out/pytest_unet.log | head -20
echo "== bench (PDL)"
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
grep -E "profiled|timed|denoise" gpurun_out/bench.err
echo "== bench (no PDL)"
UCE_NO_PDL=1 timeout 600 python bench.py --no-cpu > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "rc=$?"
grep -E "denoise" gpurun_out/bench_nopdl.err
