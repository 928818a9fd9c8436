"""Automatic traversal choice and the shadow options of the stream traversal on the bench scenes (CPTB200_LIB selects a build variant)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import cuda_pt_b200 as cpt
from stream_check import frame, timed, GOLD, VADER

tag = os.path.basename(os.environ.get("CPTB200_LIB", "default"))
S = int(os.environ.get("S", "64"))
for scene in sys.argv[1:] or [VADER]:
    r = cpt.WavefrontRenderer(scene, 0, 0)
    r.set_option("kernel_timing", 1)
    r.set_option("samples_in_flight", S)
    rows = []
    for name, trav, opts in (("auto", -1, ()), ("ordered", 2, ()), ("stream", 7, (("sq_shadow", 0),)), ("stream+shadow", 7, (("sq_shadow", 1),))):
        r.set_option("traversal", trav)
        for k, v in opts:
            r.set_option(k, v)
        r.render_batch(4 * S, False)
        ms, mr, kt = timed(r, 2 * S)
        print(json.dumps({"lib": tag, "scene": os.path.basename(scene), "mode": name, "effective": r.get_option("traversal"), "auto": r.get_option("traversal_auto"), "S": S,
                          "ms_per_spp": ms, "mrays": mr, "kernels": kt}), flush=True)
    print("   ", r.info()[:200], flush=True)
    r.close()
