#!/usr/bin/env bash
set -x
O=gpurun_out/${1:-fog}
mkdir -p $O
R=oracle/_ref/ref_dump
timeout 300 $R render scenes/xml/vader_fog_vpt_256.xml $O/vader_fog_256 vpt 0 1,2,3,4,5,6,7,8 > $O/ref.log 2>&1; tail -2 $O/ref.log
python - <<PY 2>&1 | tee $O/check.log
import os, time, numpy as np, cuda_pt_b200 as cpt
O="$O"
for trav in (1, 2):
  for pipe in (0, 3):
    r = cpt.WavefrontRenderer(O + "/vader_fog_256.cpts", 0, 0); r.set_option("pipeline", pipe); r.set_option("traversal", trav)
    for n in range(1, 9):
        r.render_raw(False)
        out = r.download_output(); ref = np.fromfile(O + f"/vader_fog_256_vpt_f{n:04d}.bin", dtype=np.float32).reshape(out.shape)
        bad = np.argwhere(~(out[..., :3] == ref[..., :3]).all(-1))
        print("trav", trav, "pipe", pipe, "frame", n, "bad", len(bad), [(int(y), int(x), out[y, x].tolist(), ref[y, x].tolist()) for y, x in bad[:2]])
    r.close()
PY
rm -f $O/*.cpts $O/*.bin
$R bench scenes/xml/vader_fog_vpt_1080p.xml vpt 8 32 2>&1 | grep ref_bench | tee $O/ref_bench.log
