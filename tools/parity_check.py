"""GPU parity check of the CUDA path against float frames written by the reference harness
(oracle/_ref/ref_dump render ...).  Usage: python tools/parity_check.py <dir with *.cpts / *_fNNNN.bin>"""
import glob
import os
import re
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt


def rel_l2(a, b):
    return float(np.linalg.norm(a[..., :3].astype(np.float64) - b[..., :3]) / np.linalg.norm(b[..., :3].astype(np.float64)))


def main(d):
    for cpts in sorted(glob.glob(os.path.join(d, "*.cpts"))):
        prefix = cpts[:-5]
        frames = sorted(glob.glob(prefix + "_wfpt_f*.bin"))
        if not frames:
            continue
        seed = 7 if prefix.endswith("_s7") else 0
        for trav in (1, 0):
            r = cpt.WavefrontRenderer(cpts, 0, seed)
            r.enable_variance(True)
            r.set_option("traversal", trav)
            W, H = r.width, r.height
            done = 0
            for f in frames:
                n = int(re.search(r"_f(\d+)\.bin$", f).group(1))
                ref = np.fromfile(f, dtype=np.float32).reshape(H, W, 4)
                while done < n:
                    r.render_raw(False)
                    done += 1
                out = r.download_output()
                if n == 1 and trav == 0 and os.environ.get("PARITY_SAVE"):
                    out.tofile(os.path.join(os.environ["PARITY_SAVE"], os.path.basename(prefix) + "_ours_f0001.bin"))
                exact = float((out[..., :3] == ref[..., :3]).all(axis=-1).mean())
                amax = float(np.abs(out[..., :3] - ref[..., :3]).max())
                a_exact = float((out[..., 3] == ref[..., 3]).mean())
                vf = f.replace("_wfpt_f", "_wfpt_var_f")
                vtxt = ""
                if os.path.exists(vf):
                    vref = np.fromfile(vf, dtype=np.float32).reshape(H, W, 1)
                    v = r.download_variance()
                    vtxt = f" var_rel={np.linalg.norm(v - vref) / max(np.linalg.norm(vref), 1e-20):.3e}"
                print(f"{os.path.basename(prefix)} trav={'dfs' if trav else 'ord'} frame {n:4d}: rel_l2={rel_l2(out, ref):.3e} "
                      f"bit-exact px={exact*100:.3f}% alpha-exact={a_exact*100:.2f}% max_abs={amax:.3e}{vtxt}", flush=True)
            # batch path must equal the per-sample path
            r2 = cpt.WavefrontRenderer(cpts, 0, seed)
            r2.set_option("traversal", trav)
            r2.render_batch(done, False)
            o2 = r2.download_output()
            print(f"    batch({done}) vs render_raw x{done}: identical={bool((o2 == out).all())} rel={rel_l2(o2, out):.2e}  stats={r2.stats()}")
            r.close(); r2.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden")
