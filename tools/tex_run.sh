python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python tools/quick_bench.py scenes/xml/vader_wfpt_1080p.xml 32 0 16
CPTB200_X=1 python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import cuda_pt_b200 as cpt
xml="scenes/xml/vader_wfpt_1080p.xml"; spp=32
for split in (0,1,0,1):
    r = cpt.WavefrontRenderer(xml, 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1); r.set_option("split_shadow", split)
    r.render_batch(32, False)
    s0 = r.stats(); k0 = r.kernel_times()
    r.render_batch(spp, False); r.sync()
    s1 = r.stats(); k1 = r.kernel_times()
    dev = (s1["device_us"] - s0["device_us"]) / 1e6
    kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
    print(f"split={split}: {dev/spp*1e3:.3f} ms/spp, {(s1['rays']-s0['rays'])/dev/1e6:.1f} Mrays/s {kt}", flush=True)
    r.close()
PY
