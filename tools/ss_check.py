"""Variant check (CPTB200_LIB): quick parity (two goldens + vader 1080p vs traversal 2), then traversal 7 at S = 16 / 64 on the given scenes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import cuda_pt_b200 as cpt
from stream_check import frame, timed, GOLD, VADER

tag = os.path.basename(os.environ.get("CPTB200_LIB", "default"))
scenes = sys.argv[1:] or [VADER]
ok = True
for scene, gold, spp in (("cbox_mats_256.cpts", "cbox_mats_256_wfpt_f0016.bin", 16), ("texroom_256.cpts", "texroom_256_wfpt_f0064.bin", 64)):
    ref = np.fromfile(os.path.join(GOLD, gold), dtype=np.float32).reshape(256, 256, 4)
    out, st, eff = frame(os.path.join(GOLD, scene), spp, 7)
    ex = float((out.view(np.uint32) == ref.view(np.uint32)).all(-1).mean())
    ok &= ex == 1.0 and eff == 7
    print(f"[{tag}] {scene} trav 7 (effective {eff}): bit-exact {ex * 100:.4f}%", flush=True)
for scene in scenes:
    a, sa, _ = frame(scene, 2, 2)
    b, sb, eff = frame(scene, 2, 7)
    ex = float((a.view(np.uint32) == b.view(np.uint32)).all(-1).mean())
    ok &= ex == 1.0 and sa["rays"] == sb["rays"]
    print(f"[{tag}] {os.path.basename(scene)} 2 spp: trav 7 (effective {eff}) vs trav 2 bit-exact {ex * 100:.5f}%  rays {sb['rays']} vs {sa['rays']}", flush=True)
print(f"[{tag}] parity " + ("OK" if ok else "FAILED"), flush=True)
for scene in scenes:
    r = cpt.WavefrontRenderer(scene, 0, 0)
    r.set_option("kernel_timing", 1)
    for trav in (2, 7):
        r.set_option("traversal", trav)
        r.set_option("sq_refill", 8), r.set_option("sq_flush", 16), r.set_option("sq_steps", 4)
        for S in (16, 64):
            r.set_option("samples_in_flight", S)
            r.render_batch(S, False)
            ms, mr, kt = timed(r, max(64, S))
            print(json.dumps({"lib": tag, "scene": os.path.basename(scene), "traversal": r.get_option("traversal"), "S": S, "ms_per_spp": ms, "mrays": mr, "kernels": kt}), flush=True)
    r.close()
