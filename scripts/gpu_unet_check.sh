#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/gpu_unet_check.py > gpurun_out/unet_check.log 2>&1
echo "exit=$?" >> gpurun_out/unet_check.log
tail -120 gpurun_out/unet_check.log
