#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm -s 150 -c 8 -o gpurun_out/prof_conv -f python scripts/profile_step.py fp16 1 > gpurun_out/ncu_conv.log 2>&1; echo "exit=$?"
tail -5 gpurun_out/ncu_conv.log
timeout 600 python -m pytest tests/test_gpu_mpm.py -m gpu -x -q 2>&1 | tail -3
