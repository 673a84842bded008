#!/bin/bash
# Round-end style GPU pass (run under gpurun from the repo root): tests, smoke, bench + reference arm, ncu launch list, ncu full captures.
# Outputs go to gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/smoke.log; grep smoke gpurun_out/smoke.log
echo "== bench (default precision fp16e5)"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit=$?"; cut -c1-2000 gpurun_out/bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "exit=$?"
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_fp16e5.csv python scripts/profile_step.py fp16e5 30 > gpurun_out/ncu_list.log 2>&1; echo "exit=$?"
echo "== ncu full: dominant conv launch (128->128 3x3x3 @ 64^3 of the warm iteration)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm -s 147 -c 1 -o gpurun_out/prof_conv -f python scripts/profile_step.py fp16e5 1 > gpurun_out/ncu_conv.log 2>&1; echo "exit=$?"
echo "== ncu full: fused MPM kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpm_fused -s 40 -c 1 -o gpurun_out/prof_mpm -f python scripts/profile_step.py fp16 30 > gpurun_out/ncu_mpm.log 2>&1; echo "exit=$?"
fi
ls -la gpurun_out | head -30
