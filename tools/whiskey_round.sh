#!/usr/bin/env bash
O=gpurun_out/whiskey; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q -k whiskey 2>&1 | tail -5
oracle/_ref/ref_dump bench scenes/ref_xml/whiskey.xml wfpt 8 32 2>&1 | grep ref_bench | tee $O/ref_bench.json
oracle/_ref/ref_dump dump scenes/ref_xml/whiskey.xml $O/w.cpts > /dev/null 2>&1
python - <<PY 2>&1 | tee $O/ours.log
import time, cuda_pt_b200 as cpt
for name, path in (("reference-built scene", "$O/w.cpts"), ("own loader", "scenes/ref_xml/whiskey.xml")):
    for trav in (1, 2):
        r = cpt.WavefrontRenderer(path, 0, 0); r.set_option("traversal", trav); r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
        r.render_batch(16); t = time.time(); r.render_batch(64); dt = time.time() - t
        st = r.stats()
        print(name, "traversal", trav, "ms/spp %.3f" % (dt / 64 * 1e3), "rays/sample %.3f" % (st["rays"] / st["samples"]), {k: round(v["ms"] / 80, 3) for k, v in r.kernel_times().items()}, flush=True)
        r.close()
PY
rm -f $O/w.cpts
