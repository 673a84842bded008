#!/bin/bash
# Bring-up of the tcgen05 conv on a B200 (run under gpurun). Logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/conv_bringup.log 2>&1
echo "=== baseline" >> gpurun_out/conv_bringup.log
timeout 300 ./build/conv_test >> gpurun_out/conv_bringup.log 2>&1
echo "exit=$?" >> gpurun_out/conv_bringup.log
if ! grep -q "gemm1x1_64_64_d16\] PASS" gpurun_out/conv_bringup.log; then
  for x in 10000 400000000000 10000400000000000; do
    echo "=== variant PIXIE_DESC_XOR=$x (gemm only)" >> gpurun_out/conv_bringup.log
    PIXIE_DESC_XOR=$x timeout 120 ./build/conv_test gemm1x1_64_64 >> gpurun_out/conv_bringup.log 2>&1
    echo "exit=$?" >> gpurun_out/conv_bringup.log
  done
fi
tail -80 gpurun_out/conv_bringup.log
