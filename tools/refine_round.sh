#!/usr/bin/env bash
O=gpurun_out/refine; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest.log
oracle/_ref/ref_dump dump scenes/ref_xml/whiskey.xml $O/w.cpts > /dev/null 2>&1
oracle/_ref/ref_dump dump scenes/xml/vader_wfpt_1080p.xml $O/v.cpts > /dev/null 2>&1
for rf in 2; do
CPTB200_REFINE_LEAVES=$rf python - <<PY 2>&1 | tee -a $O/ours.log
import time, numpy as np, cuda_pt_b200 as cpt
for name, path in (("whiskey ref-built", "$O/w.cpts"), ("vader ref-built", "$O/v.cpts"), ("vader own", "scenes/xml/vader_wfpt_1080p.xml")):
    r = cpt.WavefrontRenderer(path, 0, 0); r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
    r.render_batch(16); t = time.time(); r.render_batch(48); dt = time.time() - t
    print("refine>$rf", name, "ms/spp %.3f" % (dt / 48 * 1e3), {k: round(v["ms"] / 64, 3) for k, v in r.kernel_times().items()}, "crc", int(r.download_output().view(np.uint32).sum() % 1000003), r.info().split("pair depth")[1][:80], flush=True)
    r.close()
PY
done
rm -f $O/*.cpts
