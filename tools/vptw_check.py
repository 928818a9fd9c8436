"""GPU parity of the volumetric path, megakernel (pipeline 0) and wavefront (pipeline 3) forms, against the reference's frames."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref")
for name, frames in (("medium_cbox_256", (1, 64)), ("grid_cbox_256", (1, 16)), ("grid_cbox_rayleigh_256", (16,))):
    for pipe in (1, 3):
        for trav in (1, 2):
            r = cpt.WavefrontRenderer(os.path.join(g, name + ".cpts"), 0, 0)
            r.set_option("traversal", trav)
            r.set_option("pipeline", pipe)
            r.enable_variance(True)
            done = 0
            for n in frames:
                while done < n:
                    r.render_raw(False); done += 1
                out = r.download_output()
                ref = np.fromfile(os.path.join(g, f"{name}_vpt_f{n:04d}.bin"), dtype=np.float32).reshape(out.shape)
                ex = float((out[..., :3] == ref[..., :3]).all(-1).mean()); aex = float((out[..., 3] == ref[..., 3]).mean())
                print(f"{name} pipe={pipe} trav={trav} frame {n}: bit-exact px={ex*100:.3f}% alpha-exact={aex*100:.2f}% max_abs={np.abs(out-ref)[...,:3].max():.3e}")
                if ex < 1:
                    bad = np.argwhere(~(out[..., :3] == ref[..., :3]).all(-1))
                    for y, x in bad[:4]:
                        print("   ", y, x, out[y, x], ref[y, x])
            print("   ", r.stats())
            r.close()
for scene in ("medium_cbox_vpt_1080p.xml", "grid_cbox_vpt_1080p.xml", "vader_fog_vpt_1080p.xml"):
    path = os.path.join(os.path.dirname(g), "..", "..", "scenes", "xml", scene)
    if "grid" in scene and not os.path.exists(os.path.join(os.path.dirname(path), "..", "volume", "puff512.done")):
        continue
    for pipe in (0, 1, 3):
        r = cpt.WavefrontRenderer(path, 0, 0)
        r.set_option("pipeline", pipe)
        r.set_option("samples_in_flight", 16)
        r.set_option("kernel_timing", 1)
        r.render_batch(16)
        t = time.time(); r.render_batch(64); dt = time.time() - t
        print(scene, "pipe", pipe, "ms/spp", dt / 64 * 1e3, {k: round(v["ms"] / 80, 3) for k, v in r.kernel_times().items()}, r.stats())
        r.close()
