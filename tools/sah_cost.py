"""SAH cost of the BVH in a .cpts / XML scene (reference linear layout: 2 float4 per node, inner nodes store -(subtree size))."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cuda_pt_b200 as cpt
from cpts import read_cpts
path = sys.argv[1]
if path.endswith(".xml"):
    sc = cpt.Scene(path); path = "/tmp/_sah.cpts"; sc.save(path)
d = read_cpts(path)
n = d["nodes"].reshape(-1, 2, 4)
lo, hi = n[:, 0, :3].astype(np.float64), n[:, 1, :3].astype(np.float64)
cnt = n[:, 1, 3].copy().view(np.int32)
ext = np.maximum(hi - lo, 0)
area = 2 * (ext[:, 0] * ext[:, 1] + ext[:, 1] * ext[:, 2] + ext[:, 0] * ext[:, 2])
root = area[0]
inner = cnt < 0
print(f"nodes {len(n)}  leaves {int((~inner).sum())}  prims/leaf mean {cnt[~inner].mean():.2f} max {cnt[~inner].max()}")
print(f"SAH: inner {area[inner].sum()/root:.3f}  leaf-prim {(area[~inner]*cnt[~inner]).sum()/root:.3f}  "
      f"total(Ct=1,Ci=1) {(area[inner].sum() + (area[~inner]*cnt[~inner]).sum())/root:.3f}")
