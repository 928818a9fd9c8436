"""A/B of build variants (CPTB200_LIB=...): golden parity on three scenes, vader 1080p traversal 7 against traversal 2, then timings."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import cuda_pt_b200 as cpt
from stream_check import frame, timed, GOLD, VADER

tag = os.path.basename(os.environ.get("CPTB200_LIB", "default"))
ok = True
for scene, gold, spp in (("cbox_mats_256.cpts", "cbox_mats_256_wfpt_f0016.bin", 16), ("texroom_256.cpts", "texroom_256_wfpt_f0064.bin", 64), ("cbox_wfpt_256.cpts", "cbox_wfpt_256_wfpt_f0004.bin", 4)):
    ref = np.fromfile(os.path.join(GOLD, gold), dtype=np.float32).reshape(256, 256, 4)
    for trav, opts in ((2, ()), (7, ()), (7, (("sq_shadow", 1), ("sq_refill", 1), ("sq_flush", 1), ("sq_steps", 1))), (7, (("sq_refill", 32), ("sq_flush", 32), ("sq_steps", 16), ("sq_blocks", 1)))):
        out, st, eff = frame(os.path.join(GOLD, scene), spp, trav, opts)
        ex = float((out.view(np.uint32) == ref.view(np.uint32)).all(-1).mean())
        ok &= ex == 1.0
        print(f"[{tag}] {scene} trav {trav} {dict(opts)} (effective {eff}): bit-exact {ex * 100:.4f}%", flush=True)
a, sa, _ = frame(VADER, 4, 2)
for opts in ((), (("sq_shadow", 1),)):
    b, sb, eff = frame(VADER, 4, 7, opts)
    ex = float((a.view(np.uint32) == b.view(np.uint32)).all(-1).mean())
    ok &= ex == 1.0 and sa["rays"] == sb["rays"]
    print(f"[{tag}] vader 1080p 4 spp: trav 7 {dict(opts)} (effective {eff}) vs trav 2 bit-exact {ex * 100:.5f}%  rays {sb['rays']} vs {sa['rays']}", flush=True)
print(f"[{tag}] parity " + ("OK" if ok else "FAILED"), flush=True)
r = cpt.WavefrontRenderer(VADER, 0, 0)
r.set_option("samples_in_flight", 16)
r.set_option("kernel_timing", 1)
cfgs = [(2, ())]
for rf, fl, st in ((4, 16, 8), (8, 16, 4), (8, 16, 8), (4, 8, 4), (2, 16, 8), (16, 16, 8)):
    cfgs.append((7, (("sq_shadow", 0), ("sq_refill", rf), ("sq_flush", fl), ("sq_steps", st))))
cfgs += [(7, (("sq_shadow", 1), ("sq_refill", 4), ("sq_flush", 16), ("sq_steps", 8))), (7, (("sq_shadow", 0), ("sq_refill", 4), ("sq_flush", 16), ("sq_steps", 8), ("sq_blocks", 6))), (2, ())]
for trav, opts in cfgs:
    r.set_option("traversal", trav)
    for k, v in opts:
        r.set_option(k, v)
    r.render_batch(16, False)
    ms, mr, kt = timed(r, 48)
    print(json.dumps({"lib": tag, "traversal": trav, "opts": dict(opts), "ms_per_spp": ms, "mrays": mr, "kernels": kt}), flush=True)
    r.set_option("sq_blocks", 0)
r.close()
