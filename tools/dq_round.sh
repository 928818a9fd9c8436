#!/usr/bin/env bash
T=${1:-r02k}
O=gpurun_out/$T
mkdir -p $O
timeout 600 python tools/coop_check.py parity 2>&1 | grep -E "parity" | tee $O/parity.log
timeout 600 python tools/coop_check.py bench 2>&1 | grep -v "wide layout" | tee $O/bench.log
for RF in 4 8 16; do for FL in 8 16 24; do for ST in 2 4; do
  echo "== traversal 6: refill $RF flush $FL steps $ST" | tee -a $O/sweep.log
  CPTB200_DYN_REFILL=$RF CPTB200_DQ_FLUSH=$FL CPTB200_DYN_STEPS=$ST timeout 300 python - <<PY 2>&1 | tee -a $O/sweep.log
import sys, os, json
sys.path.insert(0, os.getcwd())
import cuda_pt_b200 as cpt
r = cpt.WavefrontRenderer("scenes/xml/vader_wfpt_1080p.xml", 0, 0)
r.set_option("traversal", 6); r.set_option("samples_in_flight", 16); r.set_option("kernel_timing", 1)
r.render_batch(32, False)
s0, k0 = r.stats(), r.kernel_times()
r.render_batch(32, False); r.sync()
s1, k1 = r.stats(), r.kernel_times()
dev = (s1["device_us"] - s0["device_us"]) / 1e6
print(json.dumps({"ms_per_spp": round(dev / 32 * 1e3, 3), "kernels": {k: round((k1[k]["ms"] - k0[k]["ms"]) / 32, 3) for k in k1}}))
r.close()
PY
done; done; done
