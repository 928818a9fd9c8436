#!/usr/bin/env bash
# GPU-box session for the cooperative traversal: parity first, then timings; optional extras.
T=${1:-r02c}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.txt
timeout 600 python tools/coop_check.py parity 2>&1 | grep -E "parity" | tee $O/parity_w4.log
timeout 900 python tools/coop_check.py bench 2>&1 | tee $O/bench_w4.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 4 --warmup 3 2>&1 | tail -1 | tee $O/bench_reference.json
timeout 900 python bench.py --steps 8 --warmup 3 2>&1 | tail -1 | tee $O/bench_ours.json
