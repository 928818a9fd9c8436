"""A/B of compile-time variants of the volumetric wavefront kernels: one subprocess per library build."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import cuda_pt_b200 as cpt
for scene in ("medium_cbox_vpt_1080p.xml", "vader_fog_vpt_1080p.xml", "grid_cbox_vpt_1080p.xml"):
    if "grid" in scene and not os.path.exists(os.path.join(%r, "scenes", "volume", "puff512.done")):
        continue
    r = cpt.WavefrontRenderer(os.path.join(%r, "scenes", "xml", scene), 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
    r.render_batch(16)
    t = time.time(); r.render_batch(48); dt = time.time() - t
    print("   ", scene, "ms/spp %%.3f" %% (dt / 48 * 1e3), {k: round(v["ms"] / 64, 3) for k, v in r.kernel_times().items()}, flush=True)
    r.close()
''' % (root, root, root)
libs = [None] + sorted(os.path.join(root, "build", "variants", f) for f in os.listdir(os.path.join(root, "build", "variants")) if f.endswith(".so"))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["CPTB200_LIB"] = lib
    print("==", os.path.basename(lib) if lib else "default", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
