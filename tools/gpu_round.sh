#!/usr/bin/env bash
# One GPU-box session: tests, smoke, both bench arms, ncu launch list + full capture of the top kernels.
# usage: bash tools/gpu_round.sh [tag]   (outputs under gpurun_out/<tag>, default r01)
set -x
T=${1:-r01}
O=gpurun_out/$T
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/smoke.log
python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | tee $O/bench_reference.json
python bench.py --steps 16 --warmup 3 2>&1 | tail -1 | tee $O/bench_ours.json
python bench.py --steps 8 --warmup 3 --traversal 1 --pipeline 0 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_ours_dfs_fused.json
python bench.py --scene c2 --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | tee $O/bench_c2_reference.json
python bench.py --scene c2 --steps 8 --warmup 3 2>&1 | tail -1 | tee $O/bench_c2_ours.json
python bench.py --scene vpt --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | tee $O/bench_vpt_reference.json
python bench.py --scene vpt --steps 8 --warmup 3 2>&1 | tail -1 | tee $O/bench_vpt_ours.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 1 -o $O/prof_trace python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shadow_apply -s 4 -c 1 -o $O/prof_shadow python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 1 -o $O/prof_shade python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la $O
