#!/usr/bin/env bash
for g in default 32 64 128; do
  echo "== L2 fetch $g"
  if [ $g != default ]; then export CPTB200_L2_FETCH=$g; fi
  python - <<PY 2>&1 | grep -v "^$"
import os, time, cuda_pt_b200 as cpt
for scene in ("medium_cbox_vpt_1080p.xml", "vader_fog_vpt_1080p.xml", "vader_wfpt_1080p.xml"):
    r = cpt.WavefrontRenderer("scenes/xml/" + scene, 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
    r.render_batch(16)
    t = time.time(); r.render_batch(48); dt = time.time() - t
    print("   ", scene, "ms/spp %.3f" % (dt / 48 * 1e3), {k: round(v["ms"] / 64, 3) for k, v in r.kernel_times().items()}, flush=True)
    r.close()
PY
done
