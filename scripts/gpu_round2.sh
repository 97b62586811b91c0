#!/bin/bash
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest SIMT" | tee -a $S
UCE_APPLY_IMPL=1 timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not tcgen05 and not impl2 and not full_model" > gpurun_out/pytest_simt.log 2>&1; echo "rc=$?" | tee -a $S
tail -15 gpurun_out/pytest_simt.log
echo "== cpu probe" | tee -a $S
timeout 300 python scripts/cpu_threads_probe.py > gpurun_out/cpu_probe.log 2>&1; echo "rc=$?" | tee -a $S
cat gpurun_out/cpu_probe.log
echo "== bench SIMT eager" | tee -a $S
timeout 300 python bench.py --apply-impl 1 --no-graph --no-cpu --steps 20 > gpurun_out/bench_simt_eager.json 2> gpurun_out/bench_simt_eager.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench_simt_eager.json; tail -12 gpurun_out/bench_simt_eager.err
echo "== bench SIMT graph" | tee -a $S
timeout 300 python bench.py --apply-impl 1 --no-cpu --steps 20 > gpurun_out/bench_simt_graph.json 2> gpurun_out/bench_simt_graph.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench_simt_graph.json; tail -12 gpurun_out/bench_simt_graph.err
echo "== pytest tcgen05" | tee -a $S
timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -x -q -p no:cacheprovider -k "tcgen05 or full_model" > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" | tee -a $S
tail -40 gpurun_out/pytest_tc.log
echo "== bench tc" | tee -a $S
timeout 300 python bench.py --apply-impl 2 --steps 20 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench_tc.json; tail -12 gpurun_out/bench_tc.err
echo "== pytest all (auto impl)" | tee -a $S
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
tail -15 gpurun_out/pytest_all.log
