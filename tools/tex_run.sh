set -x
mkdir -p gpurun_out/ref2 gpurun_out/xmlcmp
R=oracle/_ref/ref_dump
X=scenes/xml
$R render $X/texroom_wfpt_256.xml gpurun_out/ref2/texroom_256 wfpt 0 1,2,16,64 > gpurun_out/ref2/texroom.log 2>&1
tail -5 gpurun_out/ref2/texroom.log
ls -la gpurun_out/ref2
python tools/parity_check.py gpurun_out/ref2 2>&1 | tail -20
# our XML loader + PNG decoder against the reference frames
python - <<'PY'
import numpy as np, cuda_pt_b200 as cpt
r = cpt.WavefrontRenderer("scenes/xml/texroom_wfpt_256.xml", 0, 0)
for n in range(64): r.render_raw(False)
out = r.download_output()
ref = np.fromfile("gpurun_out/ref2/texroom_256_wfpt_f0064.bin", dtype=np.float32).reshape(256,256,4)
print("xml-loader vs reference f64: rel_l2 =", np.linalg.norm(out[...,:3].astype(np.float64)-ref[...,:3])/np.linalg.norm(ref[...,:3].astype(np.float64)), "mean", out[...,:3].mean(), ref[...,:3].mean())
PY
