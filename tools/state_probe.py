"""Compare per-path payload state with the reference's after `depth` bounces (debug probe).
usage: python tools/state_probe.py <xml> <workdir> <depth,depth,...>"""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
xml, work = sys.argv[1], sys.argv[2]
depths = [int(x) for x in sys.argv[3].split(",")]
os.makedirs(work, exist_ok=True)
prefix = os.path.join(work, os.path.basename(xml)[:-4])
ref = "oracle/_ref/ref_dump"
subprocess.run([ref, "render", xml, prefix, "wfpt", "0", "1"], check=True, capture_output=True)
for depth in depths:
    subprocess.run([ref, "state", xml, prefix, "0", str(depth)], check=True, capture_output=True)
    r = cpt.WavefrontRenderer(prefix + ".cpts", 0, 0)
    r.set_option("max_depth", depth); r.set_option("traversal", 1)
    r.render_raw(False)
    H, W = r.height, r.width
    print(f"== depth {depth}")
    if os.environ.get("PROBE_SAVE"):
        import shutil
        for nm in ["ray_o", "ray_d", "isect"]:
            r.debug_state(nm).tofile(os.path.join(os.environ["PROBE_SAVE"], f"ours_d{depth}_{nm}.bin"))
            shutil.copy(f"{prefix}_d{depth}_{nm}.bin", os.path.join(os.environ["PROBE_SAVE"], f"ref_d{depth}_{nm}.bin"))
        shutil.copy(prefix + ".cpts", os.path.join(os.environ["PROBE_SAVE"], "probe.cpts"))
    rng_ref = np.fromfile(f"{prefix}_d{depth}_rng.bin", dtype=np.uint32).reshape(H, W, 2)
    rng = r.debug_state("rng")
    bad_rng = (rng != rng_ref).any(-1)
    print(f"   rng state mismatch px: {bad_rng.sum()}")
    for name in ["thp", "L", "ray_o", "ray_d", "isect", "pdf"]:
        a = r.debug_state(name)
        shape = (H, W) if name == "pdf" else (H, W, 4)
        b = np.fromfile(f"{prefix}_d{depth}_{name}.bin", dtype=np.float32).reshape(shape)
        if name == "ray_d":
            tag_a, tag_b = a[..., 3].view(np.uint32), b[..., 3].view(np.uint32)
            act = (tag_b & 0x10000000) != 0
            print(f"   tag mismatch (active rays): {((tag_a != tag_b) & act).sum()}  active={act.sum()}")
            a, b = a[..., :3], b[..., :3]
        if name == "ray_o":
            ta, tb = a[..., 3], b[..., 3]
            print(f"   hit_t : bit-diff px {(ta.view(np.uint32) != tb.view(np.uint32)).sum()}  max rel {np.nanmax(np.abs(ta-tb)/np.abs(tb)):.3e}")
            a, b = a[..., :3], b[..., :3]
        if name == "isect":
            a, b = a[act], b[act]
        av, bv = a.view(np.uint32), b.view(np.uint32)
        neq = (av != bv)
        neq = neq.any(-1) if neq.ndim >= 2 and neq.shape[-1] in (3, 4) else neq
        if name == "isect":
            print(f"   isect : bit-diff active px {neq.sum()} ; per-comp diffs {(av != bv).sum(0)}")
            continue
        with np.errstate(all="ignore"):
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-20)
        rel = np.nan_to_num(rel.max(-1) if rel.ndim == 3 else rel)
        print(f"   {name:6s}: bit-diff px {neq.sum():6d} (of which rng-ok {(neq & ~bad_rng).sum():6d}), rel>1e-5: {(rel > 1e-5).sum()}, rel>1e-2: {(rel > 1e-2).sum()}")
        if name == "thp" and (neq & ~bad_rng).any():
            ys, xs = np.nonzero(neq & ~bad_rng)
            for y, x in list(zip(ys, xs))[:6]:
                print("        ", y, x, "ref", b[y, x], "ours", a[y, x])
    ys, xs = np.nonzero(bad_rng)
    for y, x in list(zip(ys, xs))[:8]:
        print("      rng-bad px", y, x, "ref", rng_ref[y, x], "ours", rng[y, x])
    r.close()
