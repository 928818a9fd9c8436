#!/usr/bin/env bash
for b in 16 24 32 40; do
CPTB200_SAH_BINS=$b python - <<PY 2>&1 | grep -v "^$"
import time, numpy as np, cuda_pt_b200 as cpt
for scene in ("scenes/xml/vader_wfpt_1080p.xml", "scenes/ref_xml/whiskey.xml"):
    r = cpt.WavefrontRenderer(scene, 0, 0); r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
    r.render_batch(16); t = time.time(); r.render_batch(48); dt = time.time() - t
    print("bins $b", scene.split("/")[-1], "ms/spp %.3f" % (dt / 48 * 1e3), {k: round(v["ms"] / 64, 3) for k, v in r.kernel_times().items()}, flush=True)
    r.close()
PY
done
