#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== solver tests" | tee -a $S
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_drivers_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED" gpurun_out/pytest_solver.log | cut -c1-200 | tee -a $S
grep -E "^E  " gpurun_out/pytest_solver.log | grep -v "where\|tensor(" | cut -c1-300 | head -10 | tee -a $S
echo "== bench cfg2" | tee -a $S
timeout 600 python bench.py --no-cpu --no-denoise > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e|Error|error" gpurun_out/bench.err | head | tee -a $S
echo "== traces" | tee -a $S
UCE_NO_OVERLAP=1 UCE_AB_TRACE=gpurun_out/ab_trace.txt UCE_CHOL_TRACE=gpurun_out/chol_trace.txt timeout 300 python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 3 --warmup 3 > /dev/null 2>&1
cat gpurun_out/chol_trace.txt | tr '\n' ';' | tee -a $S; echo | tee -a $S
grep -E " p " gpurun_out/ab_trace.txt | tee -a $S
echo "== launch list" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 6 --warmup 3 > /dev/null 2>&1
python - <<'PY' | tee -a $S
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:6]: print(r[4][:40], r[-1])
PY
