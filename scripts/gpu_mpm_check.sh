#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/gpu_mpm_check.py > gpurun_out/mpm_check.log 2>&1
echo "exit=$?" >> gpurun_out/mpm_check.log
tail -80 gpurun_out/mpm_check.log
