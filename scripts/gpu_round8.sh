#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest all gpu" | tee -a $S
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "rel err|relative error|passed|failed|Error|error" gpurun_out/pytest_all.log | head -30
echo "== bench" | tee -a $S
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench.json; tail -16 gpurun_out/bench.err
echo "== ncu launches (solver)" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu full tc kernel" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_apply_tc \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise > gpurun_out/ncu_full.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu launches (unet, full forward)" | tee -a $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches_unet.csv \
    python scripts/unet_profile.py > gpurun_out/ncu_unet.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 gpurun_out/ncu_unet.log
