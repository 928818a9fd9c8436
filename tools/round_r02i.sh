#!/usr/bin/env bash
# full GPU test-suite (incl. the new ones) + per-thread vs cooperative traversal on the other bench scenes
T=${1:-r02i}
O=gpurun_out/$T
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/pytest_gpu.log
for SC in c1 c2 whiskey; do
  for TR in 2 4; do
    timeout 600 python bench.py --scene $SC --traversal $TR --steps 6 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | tail -1 > $O/bench_${SC}_t$TR.json
    python - <<PY
import json
d=json.load(open("$O/bench_${SC}_t$TR.json"))
print("$SC traversal", d["config"]["traversal"], "ms/step", round(d["ms_per_step"],2), "Mrays/s", round(d["value"],1), {k: round(v["ms_per_step"],2) for k,v in d["stages"].items()}, "shadow", round(d["stages"]["shade"].get("of_which_shadow_kernels_ms_per_step",0),2))
PY
  done
done 2>&1 | tee $O/scenes.log
