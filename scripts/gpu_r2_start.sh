#!/bin/bash
# First GPU call of round 2: is everything still green, did the tcgen05.st register-reuse rule remove the intermittent error of the
# high-rank apply (apply impl 5 and its shared-memory variant, impl 6), how deterministic is the U-Net engine now, and the cfg4 numbers with both high-rank paths.
set -u
mkdir -p gpurun_out
S=gpurun_out/status.txt; : > $S
echo "== pytest all gpu" | tee -a $S
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed" gpurun_out/pytest_all.log
echo "== high-rank tcgen05 apply: gated parity test, 5 fresh processes" | tee -a $S
for i in 1 2 3 4 5; do
  UCE_TEST_GEMM3X=1 timeout 400 python -m pytest tests/test_solver_gpu.py -m gpu -q -p no:cacheprovider -k highrank_tcgen05 2>&1 | tail -12 | grep -E "passed|failed|FAILED" | tee -a $S
done
timeout 400 python scripts/gemm3x_diag.py > gpurun_out/gemm3x_diag.txt 2>&1; grep -E "^---|clean|wrong" gpurun_out/gemm3x_diag.txt | head -60 | tee -a $S
echo "== U-Net determinism" | tee -a $S
timeout 300 python scripts/unet_determinism_probe.py 2>&1 | tail -5 | tee -a $S
echo "== bench cfg2" | tee -a $S
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" | tee -a $S
grep -E "profiled|timed region|e2e|denoise:" gpurun_out/bench.err | tee -a $S
echo "== bench cfg4: SIMT high-rank apply vs apply impl 5" | tee -a $S
timeout 300 python bench.py --workload cfg4 --no-denoise --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; grep -E "profiled|timed region" gpurun_out/bench_cfg4.err | tee -a $S
timeout 300 python bench.py --workload cfg4 --no-denoise --no-cpu --steps 5 --warmup 3 --apply-impl 5 > gpurun_out/bench_cfg4_g3.json 2> gpurun_out/bench_cfg4_g3.err; grep -E "profiled|timed region" gpurun_out/bench_cfg4_g3.err | tee -a $S
timeout 300 python bench.py --workload cfg4 --no-denoise --no-cpu --steps 5 --warmup 3 --apply-impl 6 > gpurun_out/bench_cfg4_g3s.json 2> gpurun_out/bench_cfg4_g3s.err; grep -E "profiled|timed region" gpurun_out/bench_cfg4_g3s.err | tee -a $S
echo "== EngineGenerator (debias generation rounds on the U-Net engine; opt-in test)" | tee -a $S
UCE_TEST_ENGINE_GEN=1 timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider -k engine_generator 2>&1 | tail -3 | tee -a $S
echo "== VAE decoder engine (opt-in, never run on hardware in round 1): gated parity tests under a short timeout, then one 512x512 timing" | tee -a $S
UCE_TEST_VAE=1 timeout 600 python -m pytest tests/test_vae_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_vae.log 2>&1; echo "rc=$?" | tee -a $S
tail -15 gpurun_out/pytest_vae.log | tee -a $S
timeout 300 python scripts/vae_probe.py 2>&1 | tail -6 | tee -a $S
echo "== host path: number of copy/compute groups (default 8)" | tee -a $S
for g in 4 8 12 16 32; do
  UCE_HOST_GROUPS=$g timeout 200 python bench.py --no-denoise --no-cpu --steps 20 --warmup 5 2>&1 >/dev/null | grep -E "e2e" | sed "s/^/groups $g: /" | tee -a $S
done
echo "== plain-C consumer of the C ABI (examples/edit_host.c; opt-in test)" | tee -a $S
UCE_TEST_C_EXAMPLE=1 timeout 200 python -m pytest tests/test_cli_and_host.py -m gpu -q -p no:cacheprovider -k plain_c_consumer 2>&1 | tail -3 | tee -a $S
