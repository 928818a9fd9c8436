import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
xml = sys.argv[1]; spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
travs = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0"])]
sifs = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["8"])]
for trav in travs:
    for sif in sifs:
        r = cpt.WavefrontRenderer(xml, 0, 0)
        r.set_option("traversal", trav); r.set_option("samples_in_flight", sif); r.set_option("kernel_timing", 1)
        r.render_batch(sif * 2, False)
        s0 = r.stats(); k0 = r.kernel_times(); t0 = time.perf_counter()
        r.render_batch(spp, False); r.sync()
        dt = time.perf_counter() - t0; s1 = r.stats(); k1 = r.kernel_times()
        rays = s1["rays"] - s0["rays"]; dev = (s1["device_us"] - s0["device_us"]) / 1e6
        kt = {k: round((k1[k]["ms"] - k0[k]["ms"]) / spp, 3) for k in k1}
        print(f"{os.path.basename(xml)} trav={trav} sif={sif} bpsm={os.environ.get('CPTB200_BLOCKS_PER_SM','8')}: {dev/spp*1e3:.3f} ms/spp dev, "
              f"{rays/dev/1e6:.1f} Mrays/s, kernels ms/spp {kt}", flush=True)
        r.close()
