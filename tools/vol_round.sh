#!/usr/bin/env bash
# GPU-box session for the volumetric workloads: full GPU test suite, rays/sample of the fog scene, both bench arms
set -x
O=gpurun_out/${1:-vol}
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.log
oracle/_ref/ref_dump dump scenes/xml/vader_fog_vpt_1080p.xml $O/fog_ref.cpts 2>&1 | tail -1
python - <<PY 2>&1 | tee $O/rays.log
import json, cuda_pt_b200 as cpt
out = {}
for name, path in (("reference_scene", "$O/fog_ref.cpts"), ("own_scene", "scenes/xml/vader_fog_vpt_1080p.xml")):
    r = cpt.WavefrontRenderer(path, 0, 0)
    r.render_batch(16)
    st = r.stats(); out[name] = st["rays"] / st["samples"]; print(name, r.info(), st); r.close()
print(json.dumps(out))
PY
rm -f $O/fog_ref.cpts
for w in vpt grid fog; do
  python bench.py --scene $w --steps 6 --warmup 3 2>&1 | tail -1 | tee $O/bench_${w}_ours.json
  python bench.py --scene $w --steps 6 --warmup 3 --pipeline 0 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_${w}_ours_megakernel.json
  python bench.py --scene $w --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $O/bench_${w}_reference.json
done
