#!/bin/bash
# Round 2, call 4: K-split apply after the barrier-discipline rewrite, solve_emit factor, timestep fix
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== solver tests" | tee -a $S
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_drivers_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED" gpurun_out/pytest_solver.log | cut -c1-200 | tee -a $S
grep -E "^E  " gpurun_out/pytest_solver.log | grep -v "where\|tensor(" | cut -c1-300 | head -30 | tee -a $S
grep -c "mbarrier timeout" gpurun_out/pytest_solver.log | tee -a $S
echo "== repeat the K-split tests 3 more times (fresh processes): races show up as flaky torch.equal" | tee -a $S
for k in 1 2 3; do timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q -p no:cacheprovider -k "ksplit or cfg2_full or host_path" 2>&1 | tail -1 | tee -a $S; done
echo "== bench cfg2 (auto = K-split + overlap), no-overlap, tc3" | tee -a $S
timeout 600 python bench.py --no-cpu --no-denoise > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e|Error|error" gpurun_out/bench.err | head | tee -a $S
UCE_NO_OVERLAP=1 timeout 600 python bench.py --no-cpu --no-denoise --no-e2e > gpurun_out/bench_noov.json 2> gpurun_out/bench_noov.err; grep -E "profiled|timed region" gpurun_out/bench_noov.err | sed 's/^/no-overlap: /' | tee -a $S
timeout 600 python bench.py --no-cpu --no-denoise --no-e2e --apply-impl 4 > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; grep -E "profiled|timed region" gpurun_out/bench_tc3.err | sed 's/^/tc3: /' | tee -a $S
for w in cfg1 cfg3; do timeout 300 python bench.py --workload $w --no-cpu --no-denoise --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; grep -E "profiled|timed region" gpurun_out/bench_$w.err | sed "s/^/$w: /" | tee -a $S; done
echo "== chol trace" | tee -a $S
UCE_CHOL_TRACE=gpurun_out/chol_trace.txt timeout 300 python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 3 --warmup 3 > /dev/null 2>&1; cat gpurun_out/chol_trace.txt | tr '\n' ';' | tee -a $S; echo | tee -a $S
echo "== launch list of one step" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 6 --warmup 3 > /dev/null 2>&1
python - <<'PY' | tee -a $S
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:14]: print(r[4][:40], r[-1])
PY
echo "== unet + vae tests" | tee -a $S
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py tests/test_cli_and_host.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_unet.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED" gpurun_out/pytest_unet.log | cut -c1-200 | tee -a $S
grep -E "^E  " gpurun_out/pytest_unet.log | grep -v "where\|tensor(" | cut -c1-300 | head -10 | tee -a $S
