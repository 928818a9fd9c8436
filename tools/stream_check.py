"""Round-2 experiment driver for the stream traversal (traversal=7): parity against the reference's frames and against
traversal=2 on vader 1080p, then per-kernel timings over the runtime knobs (one renderer, options changed between batches).
usage: python tools/stream_check.py [parity|bench|sweep|one] ..."""
import json, os, sys
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
    r.close()
    return out, st, eff


def parity():
    ok = True
    for scene, gold, spp in (("cbox_mats_256.cpts", "cbox_mats_256_wfpt_f0016.bin", 16), ("texroom_256.cpts", "texroom_256_wfpt_f0064.bin", 64)):
        ref = np.fromfile(os.path.join(GOLD, gold), dtype=np.float32).reshape(256, 256, 4)
        for opts in ((), (("sq_refill", 1), ("sq_flush", 1), ("sq_steps", 1)), (("sq_refill", 32), ("sq_flush", 32), ("sq_steps", 16))):
            out, st, eff = frame(os.path.join(GOLD, scene), spp, 7, opts)
            ex = float((out.view(np.uint32) == ref.view(np.uint32)).all(-1).mean())
            print(f"[parity] {scene} trav=7 (effective {eff}) opts={opts}: bit-exact pixels vs reference frame {ex * 100:.4f}%  rays {st['rays']}", flush=True)
            ok &= ex == 1.0 and eff == 7
    a, sa, _ = frame(VADER, 4, 2)
    b, sb, eff = frame(VADER, 4, 7)
    ex = float((a.view(np.uint32) == b.view(np.uint32)).all(-1).mean())
    print(f"[parity] vader 1080p 4 spp: trav 7 (effective {eff}) vs trav 2 bit-exact pixels {ex * 100:.5f}%  rays {sb['rays']} vs {sa['rays']}", flush=True)
    ok &= ex == 1.0 and eff == 7 and sa["rays"] == sb["rays"]
    print("[parity] " + ("OK" if ok else "FAILED"), flush=True)
    return ok


def timed(r, spp):
    s0, k0 = r.stats(), r.kernel_times()
    r.render_batch(spp, False)
    r.sync()
    s1, k1 = r.stats(), r.kernel_times()
    dev = (s1["device_us"] - s0["device_us"]) / 1e6
    kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
    return round(dev / spp * 1e3, 3), round((s1["rays"] - s0["rays"]) / dev / 1e6, 1), kt


def sweep(spp=32, sif=16, full=True, scene=VADER):
    tag = os.path.basename(os.environ.get("CPTB200_LIB", "default"))
    r = cpt.WavefrontRenderer(scene, 0, 0)
    r.set_option("samples_in_flight", sif)
    r.set_option("kernel_timing", 1)
    r.set_option("traversal", 2)
    r.render_batch(sif * 2, False)
    ms, mr, kt = timed(r, spp)
    print(json.dumps({"lib": tag, "traversal": 2, "ms_per_spp": ms, "mrays": mr, "kernels": kt}), flush=True)
    r.set_option("traversal", 7)
    if r.get_option("traversal") != 7:
        print("traversal 7 not available on this scene")
        return
    r.render_batch(sif, False)
    grid = [(rf, fl, st, 0) for rf in (4, 8, 16) for fl in (8, 16, 24) for st in (2, 4, 8)] if full else [(8, 16, 4, 0), (4, 16, 8, 0), (8, 8, 8, 0), (16, 16, 4, 0)]
    best = None
    for rf, fl, st, bl in grid:
        for k, v in (("sq_refill", rf), ("sq_flush", fl), ("sq_steps", st), ("sq_blocks", bl)):
            r.set_option(k, v)
        ms, mr, kt = timed(r, spp)
        row = {"lib": tag, "traversal": 7, "refill": rf, "flush": fl, "steps": st, "blocks": bl, "ms_per_spp": ms, "mrays": mr, "kernels": kt}
        print(json.dumps(row), flush=True)
        if best is None or ms < best["ms_per_spp"]:
            best = row
    # fewer resident CTAs at the best setting (more L1 per warp)
    for bl in (6, 8):
        for k, v in (("sq_refill", best["refill"]), ("sq_flush", best["flush"]), ("sq_steps", best["steps"]), ("sq_blocks", bl)):
            r.set_option(k, v)
        ms, mr, kt = timed(r, spp)
        print(json.dumps({"lib": tag, "traversal": 7, "refill": best["refill"], "flush": best["flush"], "steps": best["steps"], "blocks": bl, "ms_per_spp": ms,
                          "mrays": mr, "kernels": kt}), flush=True)
    print("best:", json.dumps(best), flush=True)
    print("   ", r.info(), flush=True)
    r.close()


def one(trav, spp=16):
    """one batch of `spp` samples of vader 1080p (the thing to put under ncu)"""
    r = cpt.WavefrontRenderer(VADER, 0, 0)
    r.set_option("traversal", trav)
    r.set_option("samples_in_flight", spp)
    r.render_batch(spp, False)
    r.sync()
    print(r.info())
    paths, shadow = r.queue_counts()
    print("queue sizes of the batch (units of the b-th launch of the trace / shadow kernels):", json.dumps({"paths": paths, "shadow": shadow}))
    r.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if what == "one":
        one(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 16)
    elif what == "parity":
        sys.exit(0 if parity() else 1)
    elif what == "sweep":
        sweep(full=True)
    else:
        sweep(full=False)
