#!/bin/bash
mkdir -p gpurun_out
echo "== pytest unet"
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "rc=$?"
grep -E "rel err|relative error|passed|failed|Error|error|timeout" gpurun_out/pytest_unet.log | head -20
echo "== bench"
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
grep -E "profiled|timed|denoise" gpurun_out/bench.err
echo "== ncu launches (unet)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet.csv \
    python scripts/unet_profile.py > gpurun_out/ncu_unet.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/ncu_unet.log
