#!/bin/bash
# Round 2, call 8: everything once — full GPU suite, full bench line, launch lists, ncu captures, sanitizers, cfg4
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest all gpu" | tee -a $S
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rfs > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED|^SKIPPED" gpurun_out/pytest_all.log | cut -c1-200 | tee -a $S
grep -E "^E  " gpurun_out/pytest_all.log | grep -v "where\|tensor(" | cut -c1-300 | head -20 | tee -a $S
echo "== smoke" | tee -a $S
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a $S
echo "== bench (full line)" | tee -a $S
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "bench \+|Error|error" gpurun_out/bench.err | tail -14 | tee -a $S
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" | tee -a $S
for w in cfg1 cfg3 cfg4; do timeout 300 python bench.py --workload $w --no-cpu --no-denoise --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; grep -E "profiled|timed region|e2e" gpurun_out/bench_$w.err | sed "s/^/$w: /" | tee -a $S; done
echo "== traces" | tee -a $S
UCE_NO_OVERLAP=1 UCE_AB_TRACE=gpurun_out/ab_trace.txt UCE_CHOL_TRACE=gpurun_out/chol_trace.txt timeout 300 python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 3 --warmup 3 > /dev/null 2>&1
cat gpurun_out/chol_trace.txt | tr '\n' ';' | tee -a $S; echo | tee -a $S
grep -E " p " gpurun_out/ab_trace.txt | tee -a $S
echo "== launch lists" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 6 --warmup 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_profile.py > gpurun_out/unet_profile.log 2>&1
python - <<'PY' | tee -a $S
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:6]: print(r[4][:40], r[-1])
PY
echo "== ncu full" | tee -a $S
for k in apply_p_kernel apply_w_kernel solve_emit_kernel chol_small_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k -f python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 4 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | tee -a $S
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm3x_kernel -s 2 -c 2 -o gpurun_out/prof_gemm3x -f python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_gemm3x.log 2>&1; tail -1 gpurun_out/ncu_gemm3x.log | tee -a $S
echo "== compute-sanitizer" | tee -a $S
timeout 2400 bash scripts/sanitize.sh "memcheck racecheck synccheck" solver,unet,vae
cat gpurun_out/sanitize_summary.txt >> $S
