import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
xml = "scenes/xml/vader_wfpt_1080p.xml"; spp = 32
def run(pipe, env={}):
    os.environ.update({k: str(v) for k, v in env.items()})
    r = cpt.WavefrontRenderer(xml, 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1); r.set_option("pipeline", pipe)
    r.render_batch(16, False)
    s0 = r.stats(); k0 = r.kernel_times()
    r.render_batch(spp, False); r.sync()
    s1 = r.stats(); k1 = r.kernel_times()
    dev = (s1["device_us"] - s0["device_us"]) / 1e6
    kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
    print(f"pipe={pipe} {env}: {dev/spp*1e3:.3f} ms/spp {kt}", flush=True)
    r.close()
run(3)
for minb, refill, steps in [(12, 8, 8), (8, 8, 8), (12, 4, 8), (12, 16, 8), (12, 8, 32), (8, 16, 16)]:
    run(4, {"CPTB200_DYN_MINB": minb, "CPTB200_DYN_REFILL": refill, "CPTB200_DYN_STEPS": steps})
run(2, {"CPTB200_DYN_MINB": 12, "CPTB200_DYN_REFILL": 8, "CPTB200_DYN_STEPS": 8})
