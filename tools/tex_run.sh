python -m pytest tests -m gpu -q -x -k "live_reference" 2>&1 | tail -3
for w in vader_wfpt_1080p vader_tex_wfpt_1080p; do
oracle/_ref/ref_dump dump scenes/xml/$w.xml /tmp/$w.cpts > /dev/null 2>&1
python - $w <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import cuda_pt_b200 as cpt
w = sys.argv[1]
for trav in (2, 1):
    r = cpt.WavefrontRenderer(f"/tmp/{w}.cpts", 0, 0)
    r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1); r.set_option("traversal", trav)
    r.render_batch(16, False)
    s0 = r.stats(); k0 = r.kernel_times()
    r.render_batch(32, False); r.sync()
    s1 = r.stats(); k1 = r.kernel_times()
    dev = (s1["device_us"] - s0["device_us"]) / 1e6
    kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / 32, 3) for k in k1}
    print(f"{w} (reference-loaded scene) trav={trav}: {dev/32*1e3:.3f} ms/spp {kt} rays/sample {(s1['rays']-s0['rays'])/(s1['samples']-s0['samples']):.4f} | {r.info()[90:200]}", flush=True)
    if trav == 2: a = r.download_output()
    else:
        b = r.download_output()
        import numpy as np
        print("   ordered vs dfs on the SBVH scene, 48 spp: bit-identical px", float((a==b).all(-1).mean()), "rel_l2", float(np.linalg.norm(a[...,:3]-b[...,:3])/np.linalg.norm(b[...,:3])))
    r.close()
PY
done
