"""pipeline 5 (dense path records) vs pipeline 3 on the headline workloads: speed + identical frames."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for scene in ("vader_wfpt_1080p.xml", "vader_tex_wfpt_1080p.xml"):
    outs = []
    for pipe in (3, 5):
        r = cpt.WavefrontRenderer(os.path.join(root, "scenes", "xml", scene), 0, 0)
        r.set_option("pipeline", pipe); r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
        r.render_batch(16)
        t = time.time(); r.render_batch(64); dt = time.time() - t
        print(scene, "pipe", pipe, "ms/spp %.3f" % (dt / 64 * 1e3), {k: round(v["ms"] / 80, 3) for k, v in r.kernel_times().items()}, r.stats(), flush=True)
        outs.append(r.download_output()); r.close()
    print("   identical:", np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)))
