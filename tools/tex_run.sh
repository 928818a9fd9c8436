python -m pytest tests -m gpu -q -x -k "row_range or pipelines" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/r01b/bench_ours_n2.json; python -c "
import json; d=json.loads(open('gpurun_out/r01b/bench_ours_n2.json').read()); print('N=2 samples:', d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], d['config']['parallelism'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 8 --warmup 3 --split tiles 2>&1 | tail -1 > gpurun_out/r01b/bench_ours_n2_tiles.json; python -c "
import json; d=json.loads(open('gpurun_out/r01b/bench_ours_n2_tiles.json').read()); print('N=2 tiles:', d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], d['config']['parallelism'])"
