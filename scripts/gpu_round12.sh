#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "trace rc=$?"
sed -n '/CHOL/,$p' gpurun_out/trace.log
