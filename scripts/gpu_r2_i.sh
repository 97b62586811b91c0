#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== solver tests" | tee -a $S
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_drivers_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED" gpurun_out/pytest_solver.log | cut -c1-200 | tee -a $S
grep -E "^E  " gpurun_out/pytest_solver.log | grep -v "where\|tensor(" | cut -c1-300 | head -10 | tee -a $S
echo "== bench cfg4" | tee -a $S
timeout 300 python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --steps 10 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; grep -E "profiled|timed region|rror" gpurun_out/bench_cfg4.err | sed "s/^/cfg4: /" | tee -a $S
UCE_NO_OVERLAP=1 timeout 300 python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --steps 10 --warmup 3 2>&1 >/dev/null | grep -E "profiled|timed region" | sed "s/^/cfg4 no-overlap: /" | tee -a $S
UCE_GENERAL_SOLVE_GEMMS=1 timeout 300 python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --steps 10 --warmup 3 2>&1 >/dev/null | grep -E "profiled|timed region" | sed "s/^/cfg4 old solve: /" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --no-graph --steps 3 --warmup 3 > /dev/null 2>&1
python - <<'PY' | tee -a $S
import csv,collections,re
lines=[l for l in open('gpurun_out/launches_cfg4.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    k=re.sub(r"\(.*","",row["Kernel Name"])[:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(row["Metric Value"].replace(",",""))
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:10]: print(n, round(t/n/1e3,2), round(t/1e3,1), k)
PY
