#!/usr/bin/env bash
# multi-GPU bench of one or more workloads: usage ngpu_scenes.sh <N> <tag> "<scene> <scene> ..." [steps]
N=${1:-4}
T=${2:-r02n$N}
SCENES=${3:-c1}
STEPS=${4:-10}
O=gpurun_out/$T
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv | tee $O/gpus.txt
P=29700
for SC in $SCENES; do
  for SPLIT in samples tiles; do
    P=$((P+1))
    timeout 1200 $TR --master-port $P bench.py --gpus $N --scene $SC --steps $STEPS --warmup 3 --split $SPLIT 2>&1 | tail -1 > $O/bench_${SC}_ours_$SPLIT.json
  done
  P=$((P+1))
  timeout 1500 $TR --master-port $P bench.py --impl reference --gpus $N --scene $SC --steps 3 --warmup 1 2>&1 | tail -1 > $O/bench_${SC}_reference.json
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "n_gpus", d.get("n_gpus"), "value", round(d["value"],1) if d.get("value") else None, "ms/step", round(d.get("ms_per_step") or 0,2), "e2e", round((d.get("e2e") or {}).get("value") or 0,1), d.get("scaling"), "msamples/s", round(d.get("msamples_per_s") or 0,1), d.get("config",{}).get("row_bands_at_the_end"), d.get("per_rank_ms_per_frame"))
    except Exception as e:
        print(f, "unreadable", e, open(f).read()[-300:])
PY
