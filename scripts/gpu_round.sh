#!/bin/bash
# One GPU-box visit: tests, smoke, bench, ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" | tee gpurun_out/status.txt
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/status.txt
tail -30 gpurun_out/pytest_gpu.log
echo "== smoke" | tee -a gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/smoke.log
echo "== bench" | tee -a gpurun_out/status.txt
timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/status.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list" | tee -a gpurun_out/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu ${BENCH_ARGS:-} > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" | tee -a gpurun_out/status.txt
fi
${EXTRA_CMD:-true}
