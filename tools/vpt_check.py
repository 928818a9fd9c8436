"""GPU parity of the volumetric path against the reference's megakernel VPT frames (tests/golden/ref/medium_cbox_256*)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref")
for trav in (1, 2):
    r = cpt.WavefrontRenderer(os.path.join(g, "medium_cbox_256.cpts"), 0, 0)
    r.set_option("traversal", trav)
    r.enable_variance(True)
    print(r.info())
    done = 0
    for n in (1, 64):
        while done < n:
            r.render_raw(False); done += 1
        out = r.download_output()
        ref = np.fromfile(os.path.join(g, f"medium_cbox_256_vpt_f{n:04d}.bin"), dtype=np.float32).reshape(256, 256, 4)
        rel = np.linalg.norm((out - ref)[..., :3].astype(np.float64)) / np.linalg.norm(ref[..., :3].astype(np.float64))
        ex = float((out[..., :3] == ref[..., :3]).all(-1).mean()); aex = float((out[..., 3] == ref[..., 3]).mean())
        print(f"trav={trav} frame {n}: rel_l2={rel:.3e} bit-exact px={ex*100:.3f}% alpha-exact={aex*100:.2f}% max_abs={np.abs(out-ref)[...,:3].max():.3e} mean {out[...,:3].mean():.5f} vs {ref[...,:3].mean():.5f}")
        if n == 1:
            bad = np.argwhere(~(out[..., :3] == ref[..., :3]).all(-1))
            for y, x in bad[:6]:
                print("   ", y, x, out[y, x], ref[y, x])
    print(r.stats())
    r.close()
