#!/usr/bin/env bash
# final validation of the round on one B200: GPU tests, smoke, both bench arms (what the driver runs), ncu launch list
T=${1:-r02z}
O=gpurun_out/$T
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/smoke.log
timeout 1500 python bench.py --impl reference --steps 8 --warmup 3 2>&1 | tail -1 > $O/bench_reference.json
timeout 1500 python bench.py --steps 16 --warmup 3 2>&1 | tail -1 > $O/bench_ours.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/bench_under_ncu.log 2>&1
python - <<PY
import json
for n in ("reference","ours"):
    d=json.loads(open("$O/bench_%s.json"%n).read().strip().splitlines()[-1])
    print(n, "value", d.get("value"), "ms/step", d.get("ms_per_step"), "msamples/s", d.get("msamples_per_s"), "e2e", (d.get("e2e") or {}).get("value"))
    for k in ("megakernel_pt","python_serial","megakernel_rules","dropin","traced_mrays","best_msamples_per_s"):
        if k in d: print("   ", k, json.dumps(d[k])[:300])
PY
