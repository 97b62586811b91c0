#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/trace_apply.py > gpurun_out/trace.log 2>&1; echo "trace rc=$?"
head -120 gpurun_out/trace.log
