#!/bin/bash
# last check of the round: driver tests through the native artifact writer, then the gated high-rank GEMM parity test (information only)
set -u
mkdir -p gpurun_out
echo "== drivers / generate"; timeout 300 python -m pytest tests/test_drivers_gpu.py tests/test_unet_gpu.py -m gpu -q -x -p no:cacheprovider -k "drivers or generate_images or edited_weights or erase or debias" 2>&1 | tail -3 | tee gpurun_out/pytest_drivers_final.log
echo "== gemm3x (opt-in impl 5)"; UCE_TEST_GEMM3X=1 timeout 150 python -m pytest tests/test_solver_gpu.py -m gpu -q -x -p no:cacheprovider -k "highrank_tcgen05" 2>&1 | tail -25 | tee gpurun_out/pytest_gemm3x.log
