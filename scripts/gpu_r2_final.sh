#!/bin/bash
# Round 2, final call: full GPU suite, smoke, every bench line, launch lists, ncu captures of the kernels that changed, traces, sanitizers.
# Everything lands in gpurun_out/; scripts/summarize_profiles_r02.py turns it into profiles/r02_*.
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
echo "== host-buffer pipeline" | tee -a $S
timeout 300 python scripts/e2e_groups_sweep.py > gpurun_out/e2e_sweep.txt 2>&1; tail -8 gpurun_out/e2e_sweep.txt | tee -a $S
UCE_HOST_GROUPS=8 timeout 300 python scripts/e2e_trace.py 6 > gpurun_out/e2e_trace.txt 2>&1; tail -9 gpurun_out/e2e_trace.txt | tee -a $S
timeout 300 python scripts/pcie_overlap_probe.py > gpurun_out/pcie_overlap.txt 2>&1
echo "== traces" | tee -a $S
UCE_CHOL_TRACE=gpurun_out/chol_trace.txt timeout 300 python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 4 --warmup 3 > /dev/null 2>&1
cat gpurun_out/chol_trace.txt | tr '\n' ';' | tee -a $S; echo | tee -a $S
UCE_NO_OVERLAP=1 UCE_AB_TRACE=gpurun_out/ab_trace.txt timeout 300 python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 3 --warmup 3 > /dev/null 2>&1
grep -E " p " gpurun_out/ab_trace.txt | tee -a $S
echo "== launch lists" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 7 --warmup 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 220 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --no-graph --steps 3 --warmup 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet.csv python scripts/unet_profile.py > gpurun_out/unet_profile.log 2>&1
echo "== ncu full" | tee -a $S
for k in apply_p_kernel apply_w_kernel solve_emit_dmma_kernel chol_small_kernel gram_pack_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k -f python bench.py --no-cpu --no-denoise --no-e2e --no-graph --steps 4 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | tee -a $S
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_emit_dmma_kernel -s 1 -c 1 -o gpurun_out/prof_solve_dmma -f python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --no-graph --steps 2 --warmup 2 > gpurun_out/ncu_solve_dmma.log 2>&1; tail -1 gpurun_out/ncu_solve_dmma.log | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm3x_kernel -s 2 -c 2 -o gpurun_out/prof_gemm3x -f python bench.py --workload cfg4 --no-cpu --no-denoise --no-e2e --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_gemm3x.log 2>&1; tail -1 gpurun_out/ncu_gemm3x.log | tee -a $S
echo "== compute-sanitizer" | tee -a $S
timeout 2400 bash scripts/sanitize.sh "memcheck racecheck synccheck" solver,unet,vae,clip
cat gpurun_out/sanitize_summary.txt >> $S
