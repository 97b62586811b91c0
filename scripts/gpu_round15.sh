#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 530 -c 530 --csv --log-file gpurun_out/launches_unet.csv \
    python scripts/unet_profile.py > gpurun_out/ncu_unet.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/ncu_unet.log
