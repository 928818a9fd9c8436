"""Samples in flight (S) sweep on vader 1080p: ms per sample and per-kernel times for traversal 2 and 7."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cuda_pt_b200 as cpt
from stream_check import timed, VADER

scene = sys.argv[1] if len(sys.argv) > 1 else VADER
r = cpt.WavefrontRenderer(scene, 0, 0)
r.set_option("kernel_timing", 1)
for trav, opts in ((2, ()), (7, (("sq_refill", 8), ("sq_flush", 16), ("sq_steps", 4)))):
    r.set_option("traversal", trav)
    for k, v in opts:
        r.set_option(k, v)
    for S in (8, 16, 32, 64):
        r.set_option("samples_in_flight", S)
        r.render_batch(S, False)
        ms, mr, kt = timed(r, max(64, 2 * S))
        print(json.dumps({"traversal": r.get_option("traversal"), "S": S, "ms_per_spp": ms, "mrays": mr, "kernels": kt}), flush=True)
print(r.info())
r.close()
