"""GPU parity of grid media against the reference's VPT frames; usage: grid_check.py <dir> <name> [<name> ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
g = sys.argv[1]
for name in sys.argv[2:]:
    for trav in (1, 2):
        r = cpt.WavefrontRenderer(os.path.join(g, name + ".cpts"), 0, 0)
        r.set_option("traversal", trav)
        r.enable_variance(True)
        if trav == 1:
            print(r.info())
        done = 0
        for n in (1, 16):
            while done < n:
                r.render_raw(False); done += 1
            out = r.download_output()
            ref = np.fromfile(os.path.join(g, f"{name}_vpt_f{n:04d}.bin"), dtype=np.float32).reshape(out.shape)
            rel = np.linalg.norm((out - ref)[..., :3].astype(np.float64)) / np.linalg.norm(ref[..., :3].astype(np.float64))
            ex = float((out[..., :3] == ref[..., :3]).all(-1).mean()); aex = float((out[..., 3] == ref[..., 3]).mean())
            print(f"{name} trav={trav} frame {n}: rel_l2={rel:.3e} bit-exact px={ex*100:.3f}% alpha-exact={aex*100:.2f}% max_abs={np.abs(out-ref)[...,:3].max():.3e} mean {out[...,:3].mean():.5f} vs {ref[...,:3].mean():.5f}")
            if n == 1:
                bad = np.argwhere(~(out[..., :3] == ref[..., :3]).all(-1))
                for y, x in bad[:6]:
                    print("   ", y, x, out[y, x], ref[y, x])
        print(r.stats())
        r.close()
