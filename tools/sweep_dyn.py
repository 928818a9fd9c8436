import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
xml = "scenes/xml/vader_wfpt_1080p.xml"; spp = 32
def run(env={}):
    os.environ.update({k: str(v) for k, v in env.items()})
    r = cpt.WavefrontRenderer(xml, 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
    r.render_batch(16, False)
    s0 = r.stats(); k0 = r.kernel_times()
    r.render_batch(spp, False); r.sync()
    s1 = r.stats(); k1 = r.kernel_times()
    dev = (s1["device_us"] - s0["device_us"]) / 1e6
    kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
    print(f"{env}: {dev/spp*1e3:.3f} ms/spp {kt} | {r.info()[150:]}", flush=True)
    r.close()
for v in (1, 0, 1, 0):
    run({"CPTB200_L2_PERSIST": v})
