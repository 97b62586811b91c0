#!/bin/bash
# session 2, call 4: potrf quarter variants; ncu source-level capture of chol_small
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest solver" | tee -a $S
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q -x -p no:cacheprovider -k "intermediates or golden or cfg2_full or edge" > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 gpurun_out/pytest_solver.log
echo "== traces" | tee -a $S
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "rc=$?" | tee -a $S
echo "== bench" | tee -a $S
timeout 300 python bench.py --no-cpu --no-denoise > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e" gpurun_out/bench_auto.err | tee -a $S
echo "== ncu full chol_small" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_small_kernel -s 3 -c 1 -f -o gpurun_out/prof_chol_small \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu launches (solver)" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?" | tee -a $S
