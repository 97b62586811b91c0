#!/bin/bash
# session 2: apply_tc3 addend through cp.async (LDGSTS) vs TMA
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest tc3 (ldgsts addend)" | tee -a $S
UCE_TC3_ADDEND=ldgsts timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q -x -p no:cacheprovider -k "two_block or cfg2_full_model or host_path" > gpurun_out/pytest_tc3.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 gpurun_out/pytest_tc3.log
for m in tma ldgsts; do
  echo "== bench addend $m" | tee -a $S
  UCE_TC3_ADDEND=$m timeout 300 python bench.py --no-cpu --no-denoise > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "rc=$?" | tee -a $S
  grep -E "profiled|timed region|e2e" gpurun_out/bench_$m.err | tee -a $S
done
echo "== ncu full apply_tc3 (ldgsts)" | tee -a $S
UCE_TC3_ADDEND=ldgsts timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc3_kernel -s 3 -c 1 -f -o gpurun_out/prof_apply_tc3_ldgsts \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise > gpurun_out/ncu_full.log 2>&1; echo "rc=$?" | tee -a $S
