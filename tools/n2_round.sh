#!/usr/bin/env bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_n2_reference.json
python -c "
import json
for f in ('bench_n2','bench_n2_reference'):
    j=json.load(open('gpurun_out/'+f+'.json')); print(f, j.get('value'), j.get('n_gpus'), j.get('ms_per_step'), j.get('scaling'))"
