for v in _v8 _v10 _v12; do
  if [ -n "$v" ]; then export CPTB200_LIB=$PWD/cuda-pt_b200/libcpt_b200$v.so; else unset CPTB200_LIB; fi
  python - "$v" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import cuda_pt_b200 as cpt
r = cpt.WavefrontRenderer("scenes/xml/medium_cbox_vpt_1080p.xml", 0, 0)
r.set_option("samples_in_flight", 16)
r.render_batch(16, False)
s0 = r.stats()
r.render_batch(32, False); r.sync()
s1 = r.stats()
dev = (s1["device_us"] - s0["device_us"]) / 1e6
print(f"variant '{sys.argv[1]}': {dev/32*1e3:.3f} ms/spp", flush=True)
PY
done
