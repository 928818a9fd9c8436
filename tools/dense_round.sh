#!/usr/bin/env bash
python -c "
import sys; sys.argv=['bench.py']
import bench; bench.ensure_volumes()"
timeout 300 python -m pytest tests -m gpu -x -q -k "volumetric or grid_media or fog" 2>&1 | tail -3
python - <<PY 2>&1 | tee gpurun_out/vptw4.log
import os, time, cuda_pt_b200 as cpt
for scene in ("medium_cbox_vpt_1080p.xml", "grid_cbox_vpt_1080p.xml", "vader_fog_vpt_1080p.xml"):
    for pipe in (1, 3):
        r = cpt.WavefrontRenderer("scenes/xml/" + scene, 0, 0)
        r.set_option("pipeline", pipe); r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
        r.render_batch(16)
        t = time.time(); r.render_batch(48); dt = time.time() - t
        print(scene, "pipe", pipe, "ms/spp %.3f" % (dt / 48 * 1e3), {k: round(v["ms"] / 64, 3) for k, v in r.kernel_times().items()}, flush=True)
        r.close()
PY
