#!/bin/bash
# Round 2, call 2: full suite with the un-gated / new parity tests, determinism probe, compute-sanitizer over every kernel family.
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest all gpu" | tee -a $S
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rfs > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED|^SKIPPED" gpurun_out/pytest_all.log | tee -a $S
grep -E "^E  " gpurun_out/pytest_all.log | head -60 | tee -a $S
grep -E "per-tap relative|drift vs fp32" gpurun_out/pytest_all.log | tee -a $S
echo "== U-Net determinism" | tee -a $S
timeout 300 python scripts/unet_determinism_probe.py 64 2>&1 | tail -5 | tee -a $S
echo "== compute-sanitizer" | tee -a $S
timeout 2400 bash scripts/sanitize.sh "memcheck racecheck synccheck" solver,unet,vae
cat gpurun_out/sanitize_summary.txt >> $S
echo "== bench cfg2 / cfg4" | tee -a $S
timeout 600 python bench.py --no-cpu --no-denoise > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e" gpurun_out/bench.err | tee -a $S
timeout 300 python bench.py --workload cfg4 --no-denoise --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; grep -E "profiled|timed region" gpurun_out/bench_cfg4.err | tee -a $S
