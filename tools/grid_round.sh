#!/usr/bin/env bash
# GPU-box session for grid media: tests of the grid path, rays/sample on the reference-built 1080p scene, both bench arms
set -x
O=gpurun_out/${1:-grid}
mkdir -p $O
python -m pytest tests -m gpu -x -q -k "grid or media or vpt" 2>&1 | tail -5 | tee $O/pytest.log
python - <<PY 2>&1 | tee $O/prep.log
import sys; sys.argv=["bench.py"]
import bench; bench.ensure_volumes()
PY
ls -la scenes/volume
oracle/_ref/ref_dump dump scenes/xml/grid_cbox_vpt_1080p.xml $O/grid_1080p_ref.cpts 2>&1 | tail -2
python - <<PY 2>&1 | tee $O/rays.log
import json, time, cuda_pt_b200 as cpt
out = {}
for name, path in (("reference_scene", "$O/grid_1080p_ref.cpts"), ("own_scene", "scenes/xml/grid_cbox_vpt_1080p.xml")):
    t = time.time()
    r = cpt.WavefrontRenderer(path, 0, 0)
    print(name, "load s", time.time() - t, r.info())
    r.render_batch(16)
    st = r.stats()
    out[name] = st["rays"] / st["samples"]
    print(name, st)
    r.close()
print(json.dumps(out))
PY
rm -f $O/grid_1080p_ref.cpts
python bench.py --scene grid --steps 8 --warmup 3 2>&1 | tail -3 | tee $O/bench_grid_ours.json
python bench.py --scene grid --impl reference --steps 4 --warmup 1 2>&1 | tail -3 | tee $O/bench_grid_reference.json
