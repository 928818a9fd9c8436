import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_pt_b200 as cpt
xml = sys.argv[1]; spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for trav in (0, 1):
    for sif in (1, 4, 8):
        r = cpt.WavefrontRenderer(xml, 0, 0)
        r.set_option("traversal", trav); r.set_option("samples_in_flight", sif)
        r.render_batch(sif * 2, False)
        s0 = r.stats(); t0 = time.perf_counter()
        r.render_batch(spp, False); r.sync()
        dt = time.perf_counter() - t0; s1 = r.stats()
        rays = s1["rays"] - s0["rays"]; dev = (s1["device_us"] - s0["device_us"]) / 1e6
        print(f"{os.path.basename(xml)} trav={trav} sif={sif}: {dt/spp*1e3:.3f} ms/spp wall, {dev/spp*1e3:.3f} ms/spp dev, "
              f"{rays/dev/1e6:.1f} Mrays/s, rays/sample={rays/(s1['samples']-s0['samples']):.3f}", flush=True)
        r.close()
print(cpt.WavefrontRenderer(xml, 0, 0).info())
