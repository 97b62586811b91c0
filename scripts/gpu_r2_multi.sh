#!/bin/bash
# 2-GPU call: sharded edit (in-place gather) vs single GPU on cfg2 and cfg4, sharded erase test, bench under torchrun, CLIP text engine tests
set -u
mkdir -p gpurun_out
S=gpurun_out/status_multi.txt; : > $S
nvidia-smi -L | tee -a $S
echo "== CLIP text engine + sharding tests" | tee -a $S
timeout 900 python -m pytest tests/test_clip_text_gpu.py tests/test_sharding_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" | tee -a $S
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_multi.log | cut -c1-300 | head -20 | tee -a $S
echo "== bench --gpus 2 (cfg2)" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-denoise \
    > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?" | tee -a $S
python -c "
import json;d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1]);print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'sharded',d.get('sharded'))" | tee -a $S
echo "== bench --gpus 2 (cfg4)" | tee -a $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-denoise --no-e2e --workload cfg4 \
    > gpurun_out/bench_2gpu_cfg4.json 2> gpurun_out/bench_2gpu_cfg4.err; echo "rc=$?" | tee -a $S
python -c "
import json;d=json.loads(open('gpurun_out/bench_2gpu_cfg4.json').read().strip().splitlines()[-1]);print('value',d['value'],'ms',d['ms_per_step'],'sharded',d.get('sharded'))" | tee -a $S
grep -E "Error|error|Traceback" gpurun_out/bench_2gpu*.err | head | tee -a $S
