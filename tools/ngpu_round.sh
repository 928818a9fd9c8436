#!/usr/bin/env bash
# multi-GPU session: usage ngpu_round.sh <N> <tag> [extra bench scene args...]
N=${1:-2}
T=${2:-r02n$N}
O=gpurun_out/$T
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv | tee $O/gpus.txt
timeout 600 $TR --master-port 29611 tools/ddp_check.py 2>&1 | grep -E "ddp_check|Error|error" | tee $O/ddp_check.log
timeout 900 $TR --master-port 29612 bench.py --gpus $N --steps 12 --warmup 3 2>&1 | tail -1 | tee $O/bench_ours_samples.json | cut -c1-600
timeout 900 $TR --master-port 29613 bench.py --gpus $N --steps 12 --warmup 3 --split tiles 2>&1 | tail -1 | tee $O/bench_ours_tiles.json | cut -c1-300
timeout 900 $TR --master-port 29614 bench.py --gpus $N --steps 12 --warmup 3 --split tiles --no-balance 2>&1 | tail -1 | tee $O/bench_ours_tiles_static.json | cut -c1-300
timeout 1200 $TR --master-port 29615 bench.py --impl reference --gpus $N --steps 4 --warmup 2 2>&1 | tail -1 | tee $O/bench_reference.json | cut -c1-600
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "n_gpus", d.get("n_gpus"), "value", round(d["value"],1) if d.get("value") else None, "ms/step", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "scaling", d.get("scaling"), d.get("config",{}).get("row_bands_at_the_end"))
    except Exception as e:
        print(f, "unreadable", e)
PY
