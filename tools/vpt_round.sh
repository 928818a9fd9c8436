#!/usr/bin/env bash
# GPU-box session for the volumetric workload: rays/sample on the reference-built scene, then both bench arms.
set -x
T=${1:-r01v}
O=gpurun_out/$T
mkdir -p $O
oracle/_ref/ref_dump dump scenes/xml/medium_cbox_vpt_1080p.xml $O/medium_1080p_ref.cpts 2>&1 | tail -2
python - <<PY 2>&1 | tee $O/rays.log
import json, cuda_pt_b200 as cpt
out = {}
for name, path in (("reference_scene", "$O/medium_1080p_ref.cpts"), ("own_scene", "scenes/xml/medium_cbox_vpt_1080p.xml")):
    r = cpt.WavefrontRenderer(path, 0, 0)
    r.render_batch(32)
    st = r.stats()
    out[name] = st["rays"] / st["samples"]
    print(name, st)
    r.close()
print(json.dumps(out))
PY
rm -f $O/medium_1080p_ref.cpts
python bench.py --scene vpt --steps 8 --warmup 3 2>&1 | tail -3 | tee $O/bench_vpt_ours.json
python bench.py --scene vpt --impl reference --steps 4 --warmup 1 2>&1 | tail -3 | tee $O/bench_vpt_reference.json
