python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import cuda_pt_b200 as cpt
ref = np.fromfile("tests/golden/ref/cbox_mats_256_wfpt_f0016.bin", dtype=np.float32).reshape(256,256,4)[..., :3].astype(np.float64)
for scene in ("scenes/xml/cbox_mats_wfpt_256.xml", "tests/golden/ref/cbox_mats_256.cpts"):
    r = cpt.WavefrontRenderer(scene, 0, 0)
    r.render_batch(16, False)
    o = r.download_output()[..., :3].astype(np.float64); r.close()
    a, b = np.minimum(o, 4.0), np.minimum(ref, 4.0)
    bo, br = a.reshape(16,16,16,16,3).mean((1,3)), b.reshape(16,16,16,16,3).mean((1,3))
    rel = np.abs(bo-br).sum(-1)/np.maximum(br.sum(-1),1e-2)
    np.set_printoptions(linewidth=200, precision=2, suppress=True)
    print(scene, "median", np.median(rel), "p90", np.quantile(rel,0.9), "mean ratio", a.mean()/b.mean())
    print(rel)
PY
