#!/usr/bin/env bash
T=${1:-r02zz}
O=gpurun_out/$T
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 > $O/bench_ours_w3.json
python - <<PY
import json
d=json.loads(open("$O/bench_ours_w3.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["traversal"], "launches", d.get("gpu_launches"), "roofline", json.dumps(d["roofline"])[:400])
PY
