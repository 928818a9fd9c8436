#!/usr/bin/env bash
# final validation of the round on one B200: GPU tests, smoke, both bench arms (what the driver runs), ncu launch list, ncu --set full
# of the shading kernel and the two stream kernels at bounce 4 of a 64-spp batch
T=${1:-r02z}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/smoke.log
timeout 1500 python bench.py --impl reference --steps 16 --warmup 3 2>&1 | tail -1 > $O/bench_reference.json
timeout 1500 python bench.py --steps 16 --warmup 4 2>&1 | tail -1 > $O/bench_ours.json
timeout 900 python bench.py --steps 16 --warmup 4 --traversal 2 --no-extras --no-cpu-baseline 2>&1 | tail -1 > $O/bench_ours_traversal2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-extras > $O/bench_under_ncu.log 2>&1
CPTB200_SQ_SHADOW=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_traverse_stream|k_shade" --launch-skip 12 --launch-count 3 \
    -o $O/final_stream python tools/stream_check.py one 7 64 > $O/ncu_final_stream.log 2>&1
tail -n 4 $O/ncu_final_stream.log
python - <<PY
import json
for n in ("reference","ours","ours_traversal2"):
    try:
        d=json.loads(open("$O/bench_%s.json"%n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "unreadable", e); continue
    print(n, "value", d.get("value"), "ms/step", d.get("ms_per_step"), "msamples/s", d.get("msamples_per_s"), "e2e", (d.get("e2e") or {}).get("value"), (d.get("config") or {}).get("traversal"))
    for k in ("megakernel_pt","python_serial","megakernel_rules","dropin","traced_mrays","best_msamples_per_s"):
        if k in d: print("   ", k, json.dumps(d[k])[:300])
PY
