#!/bin/bash
# Round-end validation on one B200: tests, smoke, bench (both arms), probes, timelines, ncu launch lists and --set full captures.
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== gemm check" | tee -a $S
timeout 300 ./scripts/gemm_check.bin > gpurun_out/gemm_check.txt 2>&1; echo "rc=$?" | tee -a $S
tail -2 gpurun_out/gemm_check.txt | cut -c1-160
echo "== pytest all gpu" | tee -a $S
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|Error|error" gpurun_out/pytest_all.log | head -10
echo "== smoke" | tee -a $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" | tee -a $S
tail -3 gpurun_out/smoke.log
echo "== bench" | tee -a $S
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench.json; grep -E "profiled|timed|denoise|cpu" gpurun_out/bench.err | tail -8
echo "== bench reference arm" | tee -a $S
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench_ref.json
echo "== probes" | tee -a $S
timeout 120 ./scripts/fp64_probe.bin > gpurun_out/fp64_probe.txt 2>&1; echo "rc=$?" | tee -a $S
timeout 120 python scripts/copy_ceiling.py > gpurun_out/copy_ceiling.txt 2>&1; echo "rc=$?" | tee -a $S
timeout 120 python scripts/pcie_floor.py > gpurun_out/pcie_floor.txt 2>&1; echo "rc=$?" | tee -a $S
timeout 200 python scripts/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1; echo "rc=$?" | tee -a $S
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu launches (solver)" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu full apply kernels / chol_small" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc3_kernel -s 3 -c 1 -f -o gpurun_out/prof_apply_tc3 \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise > gpurun_out/ncu_full.log 2>&1; echo "rc=$?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_apply_tc2 \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise --apply-impl 3 > gpurun_out/ncu_full1.log 2>&1; echo "rc=$?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_apply_tc \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise --apply-impl 2 > gpurun_out/ncu_full1b.log 2>&1; echo "rc=$?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_small_kernel -s 3 -c 1 -f -o gpurun_out/prof_chol_small \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu --no-denoise > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu launches (unet, one forward)" | tee -a $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet.csv \
    python scripts/unet_profile.py > gpurun_out/ncu_unet.log 2>&1; echo "rc=$?" | tee -a $S
echo "== ncu full unet gemm pair / attention" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:unet_gemm_pair_kernel -s 2 -c 1 -f -o gpurun_out/prof_unet_gemm_pair \
    python scripts/unet_profile.py > gpurun_out/ncu_full3.log 2>&1; echo "rc=$?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:unet_attn_kernel -c 1 -f -o gpurun_out/prof_unet_attn \
    python scripts/unet_profile.py > gpurun_out/ncu_full4.log 2>&1; echo "rc=$?" | tee -a $S
ls -la gpurun_out/*.ncu-rep
