"""Runtime knobs of the stream traversal (closest hit + shadow rays) on vader 1080p, S = 64."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cuda_pt_b200 as cpt
from stream_check import timed, VADER
S = 64
r = cpt.WavefrontRenderer(sys.argv[1] if len(sys.argv) > 1 else VADER, 0, 0)
r.set_option("kernel_timing", 1)
r.set_option("samples_in_flight", S)
r.set_option("traversal", 7)
r.set_option("sq_shadow", 1)
r.render_batch(2 * S, False)
for rf in (4, 8, 16):
    for fl in (8, 16, 24):
        for st in (2, 4, 8):
            r.set_option("sq_refill", rf), r.set_option("sq_flush", fl), r.set_option("sq_steps", st)
            ms, mr, kt = timed(r, S)
            print(json.dumps({"refill": rf, "flush": fl, "steps": st, "ms_per_spp": ms, "trace": kt["trace"], "shadow": kt["shadow"], "shade": kt["shade"]}), flush=True)
r.close()
