#!/bin/bash
# Round 2, call 3: the K-split apply (apply_ab.cu) + fused edit call; drift probe; racecheck classified by kernel
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== K-split apply tests" | tee -a $S
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q -p no:cacheprovider -x -k "ksplit or cfg2_full or host_path or golden" > gpurun_out/pytest_ab.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_ab.log | head -30 | tee -a $S
echo "== bench cfg2 (auto = K-split), tc3 for comparison" | tee -a $S
timeout 600 python bench.py --no-cpu --no-denoise > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e|Error|error" gpurun_out/bench.err | tee -a $S
UCE_NO_OVERLAP=1 timeout 600 python bench.py --no-cpu --no-denoise --no-e2e > gpurun_out/bench_noov.json 2> gpurun_out/bench_noov.err; grep -E "profiled|timed region" gpurun_out/bench_noov.err | sed 's/^/no-overlap: /' | tee -a $S
timeout 600 python bench.py --no-cpu --no-denoise --no-e2e --apply-impl 4 > gpurun_out/bench_tc3.json 2> gpurun_out/bench_tc3.err; grep -E "profiled|timed region" gpurun_out/bench_tc3.err | sed 's/^/tc3: /' | tee -a $S
for w in cfg1 cfg3; do timeout 300 python bench.py --workload $w --no-cpu --no-denoise --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; grep -E "profiled|timed region" gpurun_out/bench_$w.err | sed "s/^/$w: /" | tee -a $S; done
echo "== full suite" | tee -a $S
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rfs > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED|^SKIPPED" gpurun_out/pytest_all.log | tee -a $S
grep -E "^E  " gpurun_out/pytest_all.log | cut -c1-300 | head -40 | tee -a $S
echo "== drift probe" | tee -a $S
timeout 600 python scripts/drift_probe.py 2>&1 | tail -16 | tee -a $S
echo "== racecheck, classified" | tee -a $S
CS=/usr/local/cuda/bin/compute-sanitizer
for t in solver unet vae; do
  timeout 900 $CS --tool racecheck --racecheck-report all --print-limit 100000 python tests/tools/sanitize_target.py $t > gpurun_out/racecheck_$t.log 2>&1
  echo "-- $t: $(grep -E 'RACECHECK SUMMARY' gpurun_out/racecheck_$t.log)" | tee -a $S
  grep -E "Error: |Warning: " -A2 gpurun_out/racecheck_$t.log | grep -E " at " | sed -E 's/.* at ([A-Za-z_:0-9]+)\(.*\)\+0x[0-9a-f]+( in )?(.*)/\1 \3/' | sort | uniq -c | sort -rn | head -20 | tee -a $S
done
echo "== memcheck on the new kernels" | tee -a $S
timeout 600 $CS --tool memcheck python tests/tools/sanitize_target.py solver > gpurun_out/memcheck_solver.log 2>&1; grep -E "ERROR SUMMARY|solver \(" gpurun_out/memcheck_solver.log | cut -c1-200 | tee -a $S
