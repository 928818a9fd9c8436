python -m pytest tests -m gpu -q -x -k "pipelines or pt_rules or megakernel" 2>&1 | tail -15
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import cuda_pt_b200 as cpt
xml="scenes/xml/vader_wfpt_1080p.xml"; spp=32
for pipe in (1,3,1,3):
    r = cpt.WavefrontRenderer(xml, 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1); r.set_option("pipeline", pipe)
    r.render_batch(32, False)
    s0 = r.stats(); k0 = r.kernel_times()
    r.render_batch(spp, False); r.sync()
    s1 = r.stats(); k1 = r.kernel_times()
    dev = (s1["device_us"] - s0["device_us"]) / 1e6
    kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
    print(f"pipeline={pipe}: traced {(s1["shadow_traced"]-s0["shadow_traced"])/max(1,(s1["shadow_rays"]-s0["shadow_rays"])):.3f} of shadow rays; {dev/spp*1e3:.3f} ms/spp, {(s1['rays']-s0['rays'])/dev/1e6:.1f} Mrays/s {kt}", flush=True)
    r.close()
PY
