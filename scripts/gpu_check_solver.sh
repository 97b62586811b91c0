#!/bin/bash
# session 2: pre-final check: all solver + driver tests, traces, bench (auto impl)
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest solver+drivers" | tee -a $S
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_drivers_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 gpurun_out/pytest_solver.log
echo "== traces" | tee -a $S
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "rc=$?" | tee -a $S
echo "== bench" | tee -a $S
timeout 300 python bench.py --no-cpu --no-denoise > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e" gpurun_out/bench_auto.err | tee -a $S
cat gpurun_out/bench_auto.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['roofline'])"
