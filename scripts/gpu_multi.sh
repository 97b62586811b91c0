#!/bin/bash
# 2-GPU validation: the driver's launch line for bench.py, the layer-sharded solve + all-gather, and the erase CLI path under torchrun.
set -u
mkdir -p gpurun_out
S=gpurun_out/status_multi.txt; : > $S
nvidia-smi -L | tee -a $S
echo "== bench --gpus 2" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
    > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench_2gpu.json; grep -E "bench \+|Error|error" gpurun_out/bench_2gpu.err | tail -20
echo "== reference arm under torchrun" | tee -a $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
    > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "rc=$?" | tee -a $S
cat gpurun_out/bench_ref_2gpu.json
echo "== sharded erase driver (NCCL all-gather) vs single GPU" | tee -a $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/tools/sharded_erase_check.py \
    > gpurun_out/sharded_check.log 2>&1; echo "rc=$?" | tee -a $S
tail -5 gpurun_out/sharded_check.log
