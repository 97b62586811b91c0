#!/bin/bash
# session 2, call 5: apply_tc2 with per-unit Qt barrier, TMA-warp stores, 5-deep raw ring
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest tc2" | tee -a $S
timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q -x -p no:cacheprovider -k "two_cta or cfg2_full_model or host_path or linearity" > gpurun_out/pytest_tc2.log 2>&1; echo "rc=$?" | tee -a $S
tail -6 gpurun_out/pytest_tc2.log
echo "== traces" | tee -a $S
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "rc=$?" | tee -a $S
for v in "3 128" "3 96" "3 112"; do
  set -- $v
  echo "== bench impl $1 tile_rows $2" | tee -a $S
  UCE_TC2_TILE_ROWS=$2 timeout 300 python bench.py --no-cpu --no-denoise --apply-impl $1 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "rc=$?" | tee -a $S
  grep -E "profiled|timed region|e2e" gpurun_out/bench_$1_$2.err | tee -a $S
done
echo "== ncu full apply_tc2" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_apply_tc2 \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise --apply-impl 3 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?" | tee -a $S
