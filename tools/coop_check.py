"""Round-2 experiment driver for the cooperative traversal (traversal=4): parity against traversal=2 / the reference's
frames, then per-kernel timings over the knobs.  usage: python tools/coop_check.py [parity|bench] ..."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cuda_pt_b200 as cpt

GOLD = os.path.join(ROOT, "tests", "golden", "ref")
VADER = os.path.join(ROOT, "scenes", "xml", "vader_wfpt_1080p.xml")


def frame(scene, spp, trav, opts=()):
    r = cpt.WavefrontRenderer(scene, 0, 0)
    r.set_option("traversal", trav)
    for k, v in opts:
        r.set_option(k, v)
    eff = r.get_option("traversal")
    out = r.render_to_host(spp, False)
    st = r.stats()
    info = r.info()
    r.close()
    return out, st, eff, info


def parity():
    ok = True
    for scene, gold, spp in (("cbox_wfpt_256.cpts", "cbox_wfpt_256_wfpt_f0004.bin", 4), ("cbox_mats_256.cpts", "cbox_mats_256_wfpt_f0016.bin", 16),
                             ("texroom_256.cpts", "texroom_256_wfpt_f0064.bin", 64)):
        ref = np.fromfile(os.path.join(GOLD, gold), dtype=np.float32).reshape(256, 256, 4)
        for opts in ((), (("coop_slab", 16),)):
            out, st, eff, info = frame(os.path.join(GOLD, scene), spp, 4, opts)
            ex = float((out.view(np.uint32) == ref.view(np.uint32)).all(-1).mean())
            print(f"[parity] {scene} trav=4 (effective {eff}) opts={opts}: bit-exact pixels vs reference frame {ex * 100:.4f}%  rays {st['rays']}", flush=True)
            ok &= ex == 1.0
        print("   ", info, flush=True)
    if os.path.exists(VADER):
        a, sa, _, _ = frame(VADER, 4, 2)
        for tr in (4, 5, 6):
            b, sb, eff, info = frame(VADER, 4, tr)
            ex = float((a.view(np.uint32) == b.view(np.uint32)).all(-1).mean())
            print(f"[parity] vader 1080p 4 spp: trav {tr} (effective {eff}) vs trav 2 bit-exact pixels {ex * 100:.5f}%  rays {sb['rays']} vs {sa['rays']}", flush=True)
            ok &= ex > 0.99999 and eff == tr
        print("   ", info, flush=True)
        for scene, gold, spp in (("cbox_mats_256.cpts", "cbox_mats_256_wfpt_f0016.bin", 16), ("texroom_256.cpts", "texroom_256_wfpt_f0064.bin", 64)):
            ref = np.fromfile(os.path.join(GOLD, gold), dtype=np.float32).reshape(256, 256, 4)
            for tr in (5, 6):
                out, st, eff, info = frame(os.path.join(GOLD, scene), spp, tr)
                ex = float((out.view(np.uint32) == ref.view(np.uint32)).all(-1).mean())
                print(f"[parity] {scene} trav={tr} (effective {eff}): bit-exact pixels vs reference frame {ex * 100:.4f}%", flush=True)
                ok &= ex == 1.0 and eff == tr
    print("[parity] " + ("OK" if ok else "FAILED"), flush=True)
    return ok


def bench(spp=32, sif=16):
    rows = []
    configs = [(2, ()), (4, (("coop_slab", 0),)), (5, ()), (6, ())]
    for trav, opts in configs:
        r = cpt.WavefrontRenderer(VADER, 0, 0)
        r.set_option("traversal", trav)
        for k, v in opts:
            r.set_option(k, v)
        r.set_option("samples_in_flight", sif)
        r.set_option("kernel_timing", 1)
        r.render_batch(sif * 2, False)
        s0, k0 = r.stats(), r.kernel_times()
        r.render_batch(spp, False)
        r.sync()
        s1, k1 = r.stats(), r.kernel_times()
        dev = (s1["device_us"] - s0["device_us"]) / 1e6
        kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
        row = {"width": os.environ.get("CPTB200_COOP_WIDTH", "4"), "traversal": r.get_option("traversal"), "opts": dict(opts), "ms_per_spp": round(dev / spp * 1e3, 3),
               "mrays": round((s1["rays"] - s0["rays"]) / dev / 1e6, 1), "kernels_ms_per_spp": kt}
        print(json.dumps(row), flush=True)
        if trav == 4 and not rows:
            print("   ", r.info(), flush=True)
        rows.append(row)
        r.close()
    return rows


def one(trav, spp=16):
    """one batch of `spp` samples of vader 1080p (the thing to put under ncu)"""
    r = cpt.WavefrontRenderer(VADER, 0, 0)
    r.set_option("traversal", trav)
    r.set_option("samples_in_flight", spp)
    r.render_batch(spp, False)
    r.sync()
    print(r.info())
    paths, shadow = r.queue_counts()
    print("queue sizes of the batch (units of the b-th launch of k_trace / k_shadow_apply):", json.dumps({"paths": paths, "shadow": shadow}))
    r.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if what == "one":
        one(int(sys.argv[2]))
        sys.exit(0)
    if what == "parity":
        sys.exit(0 if parity() else 1)
    bench()
