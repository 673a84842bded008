#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./build/conv_test > gpurun_out/conv_test.log 2>&1; grep -E "TIME|FAIL|SUMMARY" gpurun_out/conv_test.log | tail -12
timeout 1200 python -m pytest tests/test_gpu_mpm.py -m gpu -x -q > gpurun_out/pytest_mpm.log 2>&1; echo "pytest mpm exit=$?"; tail -12 gpurun_out/pytest_mpm.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --precision fp16 --skip-cpu > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_fp16.err
python - <<'PY'
import json
for f in ['bench_fp16.json']:
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, 'voxels/s=%.3e'%d['value'], 'unet_ms=%.2f'%d['unet_ms_per_scene'], 'mpm us/substep=%.2f'%d['mpm']['us_per_substep'], 'conv frac=%.3f'%d['roofline']['frac'], 'mpm frac=%.3f'%d['roofline_mpm']['frac'], d['unet_kernel_breakdown_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpm_tiled -s 20 -c 2 -o gpurun_out/prof_mpm_tiled -f python scripts/profile_step.py fp16 60 > gpurun_out/ncu_mpm2.log 2>&1; echo "ncu exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm -s 150 -c 4 -o gpurun_out/prof_conv3 -f python scripts/profile_step.py fp16 1 > gpurun_out/ncu_conv3.log 2>&1; echo "ncu exit=$?"
