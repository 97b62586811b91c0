#!/bin/bash
# session 2: factor check (parity of intermediates, phase trace, bench) + flat D2D copy of the cfg2 footprint for context
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest factor" | tee -a $S
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q -x -p no:cacheprovider -k "intermediates or golden or cfg2_full or edge or primal or dense" > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?" | tee -a $S
tail -4 gpurun_out/pytest_solver.log
echo "== traces" | tee -a $S
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "rc=$?" | tee -a $S
echo "== bench" | tee -a $S
timeout 300 python bench.py --no-cpu --no-denoise > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e" gpurun_out/bench_auto.err | tee -a $S
echo "== copy ceiling" | tee -a $S
timeout 300 python scripts/copy_ceiling.py 2>&1 | tee gpurun_out/copy_ceiling.txt
