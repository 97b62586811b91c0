#!/bin/bash
# session 2, call 7: apply_tc3 with one MMA warp per row block
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest tc3" | tee -a $S
timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q -x -p no:cacheprovider -k "two_block or cfg2_full_model or host_path or linearity" > gpurun_out/pytest_tc3.log 2>&1; echo "rc=$?" | tee -a $S
tail -6 gpurun_out/pytest_tc3.log
echo "== traces" | tee -a $S
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "rc=$?" | tee -a $S
for v in "4 0" "4 128"; do
  set -- $v
  echo "== bench impl $1 block_rows $2" | tee -a $S
  if [ "$2" = "0" ]; then unset UCE_TC3_BLOCK_ROWS; else export UCE_TC3_BLOCK_ROWS=$2; fi
  timeout 300 python bench.py --no-cpu --no-denoise --apply-impl $1 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "rc=$?" | tee -a $S
  grep -E "profiled|timed region|e2e" gpurun_out/bench_$1_$2.err | tee -a $S
done
unset UCE_TC3_BLOCK_ROWS
echo "== ncu full apply_tc3" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc3_kernel -s 3 -c 1 -f -o gpurun_out/prof_apply_tc3 \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise --apply-impl 4 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?" | tee -a $S
