#!/usr/bin/env bash
# One GPU-box session: tests, smoke, both bench arms, ncu launch list + full capture of the top kernels.
set -x
mkdir -p gpurun_out/r01
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r01/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r01/smoke.log
python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r01/bench_reference.json
python bench.py --steps 16 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r01/bench_ours.json
python bench.py --steps 8 --warmup 3 --traversal 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r01/bench_ours_dfs.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 2 -o gpurun_out/r01/prof_trace python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 2 -o gpurun_out/r01/prof_shade python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r01
