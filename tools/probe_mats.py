"""Debug probe: which materials do the paths that differ from the reference after `depth` bounces go through?
usage: python tools/probe_mats.py <xml> <workdir> <depth>"""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
import cpts as cptsio
xml, work, depth = sys.argv[1], sys.argv[2], int(sys.argv[3])
os.makedirs(work, exist_ok=True)
prefix = os.path.join(work, os.path.basename(xml)[:-4])
ref = "oracle/_ref/ref_dump"
subprocess.run([ref, "render", xml, prefix, "wfpt", "0", "1"], check=True, capture_output=True)
d = cptsio.read_cpts(prefix + ".cpts")
obj_idx = d["obj_idxs"].view(np.int32).ravel()
objs = np.frombuffer(d["objects"].tobytes(), dtype=np.dtype([("be", "<i4"), ("po", "<i4"), ("pn", "<i4"), ("ia", "<f4")]))
btype = np.frombuffer(d["bsdfs"].tobytes(), dtype=np.int32).reshape(len(d["bsdfs"]), -1)[:, 0]
states = {}
for dd in (depth - 1, depth):
    if dd < 1:
        continue
    subprocess.run([ref, "state", xml, prefix, "0", str(dd)], check=True, capture_output=True)
    r = cpt.WavefrontRenderer(prefix + ".cpts", 0, 0)
    r.set_option("max_depth", dd); r.set_option("traversal", 1)
    r.render_raw(False)
    H, W = r.height, r.width
    st = {}
    for name in ["thp", "L", "ray_o", "ray_d", "isect", "pdf", "rng"]:
        a = r.debug_state(name)
        dt = np.uint32 if name == "rng" else np.float32
        shape = (H, W) if name == "pdf" else ((H, W, 2) if name == "rng" else (H, W, 4))
        b = np.fromfile(f"{prefix}_d{dd}_{name}.bin", dtype=dt).reshape(shape)
        st[name] = (a, b)
    states[dd] = st
    r.close()
st = states[depth]
a, b = st["thp"]
La, Lb = st["L"]
bad = (a.view(np.uint32) != b.view(np.uint32)).any(-1) | (La.view(np.uint32) != Lb.view(np.uint32)).any(-1)
ok_rng = (st["rng"][0] == st["rng"][1]).all(-1)
print("depth", depth, "thp/L mismatch px", bad.sum(), "with rng ok", (bad & ok_rng).sum())
if depth - 1 in states:
    pa, pb = states[depth - 1]["thp"]
    pLa, pLb = states[depth - 1]["L"]
    prev_ok = (pa.view(np.uint32) == pb.view(np.uint32)).all(-1) & (pLa.view(np.uint32) == pLb.view(np.uint32)).all(-1)
    tag_prev = states[depth - 1]["ray_d"][1][..., 3].view(np.uint32)
else:
    prev_ok = np.ones_like(bad)
    tag_prev = None
new_bad = bad & prev_ok
print("newly bad at this depth:", new_bad.sum())
ys, xs = np.nonzero(new_bad)
from collections import Counter
cnt = Counter()
for y, x in zip(ys, xs):
    # the material shaded at bounce `depth-1` is the one hit by the ray stored after depth-1 bounces
    if tag_prev is None:
        continue
    t = tag_prev[y, x]
    prim = int(t & 0x0fffffff)
    o = int(obj_idx[prim]) & 0xfffff
    bs = int(objs["be"][o]) >> 16
    cnt[(o, bs, int(btype[bs]))] += 1
print("(obj, bsdf, type) histogram of shaded hits for newly-bad paths:", cnt.most_common())
for y, x in list(zip(ys, xs))[:8]:
    print(y, x, "thp ref", b[y, x], "ours", a[y, x], "| L ref", Lb[y, x], "ours", La[y, x])
