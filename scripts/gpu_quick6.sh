#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./build/conv_test > gpurun_out/conv_test.log 2>&1; grep -E "TIME|FAIL|SUMMARY|PLAN" gpurun_out/conv_test.log | tail -16
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/pytest_unet.log 2>&1; echo "pytest unet exit=$?"; tail -3 gpurun_out/pytest_unet.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision fp16 --skip-cpu > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_fp16.err
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_fp16x3.json 2> gpurun_out/bench_fp16x3.err; echo "bench exit=$?"
python - <<'PY'
import json
for f in ['bench_fp16.json','bench_fp16x3.json']:
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, 'voxels/s=%.3e'%d['value'], 'unet_ms=%.2f'%d['unet_ms_per_scene'], 'mpm us/substep=%.2f'%d['mpm']['us_per_substep'], 'conv frac=%.3f'%d['roofline']['frac'], d['unet_kernel_breakdown_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
