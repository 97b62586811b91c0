#!/bin/bash
# session 2, call 2: elect-pattern issue in every tcgen05 kernel, rolled potrf + restructured chol_small, deeper tc2 epilogue
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest solver" | tee -a $S
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_drivers_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?" | tee -a $S
tail -8 gpurun_out/pytest_solver.log
echo "== traces" | tee -a $S
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "rc=$?" | tee -a $S
for v in "2 128" "3 128" "3 96"; do
  set -- $v
  echo "== bench impl $1 tile_rows $2" | tee -a $S
  UCE_TC2_TILE_ROWS=$2 timeout 300 python bench.py --no-cpu --no-denoise --apply-impl $1 > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "rc=$?" | tee -a $S
  grep -E "profiled|timed region|e2e" gpurun_out/bench_$1_$2.err | tee -a $S
done
echo "== pytest unet" | tee -a $S
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 gpurun_out/pytest_unet.log
echo "== bench full (auto impl, denoise)" | tee -a $S
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e|denoise" gpurun_out/bench.err | tee -a $S
