import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
xml = "scenes/xml/vader_wfpt_1080p.xml"; spp = 32
def run(pipe, trav, env={}):
    os.environ.update({k: str(v) for k, v in env.items()})
    r = cpt.WavefrontRenderer(xml, 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1); r.set_option("pipeline", pipe); r.set_option("traversal", trav)
    r.render_batch(16, False)
    s0 = r.stats(); k0 = r.kernel_times()
    r.render_batch(spp, False); r.sync()
    s1 = r.stats(); k1 = r.kernel_times()
    dev = (s1["device_us"] - s0["device_us"]) / 1e6
    kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
    print(f"pipe={pipe} trav={trav} {env}: {dev/spp*1e3:.3f} ms/spp {kt} | {r.info()[40:140]}", flush=True)
    r.close()
for t in (0.2, 0.5, 1.0, 2.0, 4.0):
    run(3, 2, {"CPTB200_SAH_TRAVERSE": t})
