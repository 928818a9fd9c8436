"""Debug helper: render variants of the medium scene with the reference harness and with our kernels on the same box."""
import os, re, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "scenes/xml/medium_cbox_vpt_256.xml")).read()
out = os.path.join(root, "gpurun_out", "vptv"); os.makedirs(out, exist_ok=True)
def variant(name, text):
    p = os.path.join(root, "scenes", "xml", f"_v_{name}.xml")
    open(p, "w").write(text)
    return p
def strip_medium(text, mesh):
    # remove the medium binding of one shape
    pat = re.compile(r'(<string name="filename" value="\.\./meshes/cbox/%s"/>\s*<ref type="material" id="[^"]+"/>)\s*<ref type="medium" id="[^"]+"/>\s*<bool name="cullable" value="true"/>' % re.escape(mesh))
    return pat.sub(r'\1', text)
def bounces(text, n):
    for k in ("max_bounce", "max_diffuse", "max_specular", "max_transmit", "max_volume"):
        text = re.sub(r'(<integer name="%s"\s+value=")\d+(")' % k, r'\g<1>%d\2' % n, text)
    return text
V = {}
V["all_b1"] = bounces(src, 1)
V["all_b2"] = bounces(src, 2)
only_tall = strip_medium(strip_medium(src, "bunny.obj"), "short-box.obj")
only_bunny = strip_medium(strip_medium(src, "tall-box.obj"), "short-box.obj")
only_short = strip_medium(strip_medium(src, "tall-box.obj"), "bunny.obj")
V["tall_hg"] = only_tall; V["bunny_iso"] = only_bunny; V["short_duo"] = only_short
V["tall_hg_b2"] = bounces(only_tall, 2); V["tall_hg_b3"] = bounces(only_tall, 3)
none = strip_medium(only_tall, "tall-box.obj")
tall_iso = only_tall.replace('<phase type="hg" id="large-vol">', '<phase type="isotropic" id="large-vol">')
tall_thin = only_tall.replace('<float name="scale" value="20"/>', '<float name="scale" value="0.5"/>')
V = {"none_b2": bounces(strip_medium(only_tall, "tall-box.obj"), 2), "all_b10": src}
for name, text in V.items():
    xml = variant(name, text)
    pre = os.path.join(out, name)
    subprocess.run([os.path.join(root, "oracle/_ref/ref_dump"), "render", xml, pre, "vpt", "0", "1"], capture_output=True, cwd=root)
    ref = np.fromfile(pre + "_vpt_f0001.bin", dtype=np.float32).reshape(256, 256, 4)
    r = cpt.WavefrontRenderer(pre + ".cpts", 0, 0)
    r.render_raw(False)
    o = r.download_output(); r.close()
    o.tofile(pre + "_ours_f0001.bin")
    ex = (o[..., :3] == ref[..., :3]).all(-1)
    d = np.abs(o[..., :3] - ref[..., :3]).max(-1)
    print(f"{name:12s}: bit-exact px {ex.mean()*100:7.3f}%  alpha-exact {(o[...,3]==ref[...,3]).mean()*100:7.3f}%  small-diff px (<1e-4 rel) {((~ex) & (d < 1e-4*np.maximum(ref[...,:3].max(-1),1e-3))).mean()*100:6.3f}%  big-diff px {((~ex) & (d >= 1e-4*np.maximum(ref[...,:3].max(-1),1e-3))).mean()*100:6.3f}%", flush=True)
    os.remove(xml)

