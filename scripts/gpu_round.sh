#!/bin/bash
# Round-end style GPU pass: tests, smoke, bench (both precisions), ncu launch list, ncu full capture.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/smoke.log; grep smoke gpurun_out/smoke.log
echo "== bench fp16x3" ; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_fp16x3.json 2> gpurun_out/bench_fp16x3.err; echo "exit=$?"; cat gpurun_out/bench_fp16x3.json | cut -c1-3000
echo "== bench fp16" ; timeout 900 python bench.py --steps 5 --warmup 3 --precision fp16 --skip-cpu > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "exit=$?"; cat gpurun_out/bench_fp16.json | cut -c1-2000
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_fp16x3.csv python scripts/profile_step.py fp16x3 30 > gpurun_out/ncu_list.log 2>&1; echo "exit=$?"
echo "== ncu full: conv kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm -s 330 -c 4 -o gpurun_out/prof_conv -f python scripts/profile_step.py fp16 1 > gpurun_out/ncu_conv.log 2>&1; echo "exit=$?"
echo "== ncu full: mpm kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpm_ -s 30 -c 3 -o gpurun_out/prof_mpm -f python scripts/profile_step.py fp16 20 > gpurun_out/ncu_mpm.log 2>&1; echo "exit=$?"
fi
ls -la gpurun_out | head -30
