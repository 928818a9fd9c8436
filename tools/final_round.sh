#!/usr/bin/env bash
set -x
O=gpurun_out/${1:-final}
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/smoke.log
python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json
python bench.py --steps 16 --warmup 3 2>&1 | tail -1 > $O/bench_ours.json
python bench.py --scene whiskey --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $O/bench_whiskey_reference.json
python bench.py --scene whiskey --steps 8 --warmup 3 2>&1 | tail -1 > $O/bench_whiskey_ours.json
for w in vpt grid fog; do
  python bench.py --scene $w --steps 6 --warmup 3 2>&1 | tail -1 > $O/bench_${w}_ours.json
done
python - <<PY
import json
for f in ("bench_reference","bench_ours","bench_whiskey_reference","bench_whiskey_ours","bench_vpt_ours","bench_grid_ours","bench_fog_ours"):
    j=json.loads(open("$O/"+f+".json").read().strip().splitlines()[-1]); print(f, j["value"], j["ms_per_step"], j.get("e2e",{}).get("value"), j.get("roofline",{}).get("kernel"), j.get("gpu_launches"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_vpt.csv python bench.py --scene vpt --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_vpt_under_ncu.log 2>&1
