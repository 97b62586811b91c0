#!/bin/bash
mkdir -p gpurun_out
echo "== pytest solver"
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_drivers_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/pytest_solver.log
echo "== bench"
timeout 300 python bench.py --no-denoise --no-cpu > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "rc=$?"
grep -E "profiled|timed" gpurun_out/bench_s.err
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "trace rc=$?"
sed -n '/CHOL/,$p' gpurun_out/trace.log
