#!/usr/bin/env bash
T=${1:-r02d}
O=gpurun_out/$T
mkdir -p $O
timeout 600 python tools/coop_check.py parity 2>&1 | grep -E "parity" | tee $O/parity.log
timeout 900 python tools/coop_check.py bench 2>&1 | tee $O/bench.log
