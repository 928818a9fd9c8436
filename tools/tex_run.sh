for v in "" _mb8 _mb5 "" _mb8; do
  if [ -n "$v" ]; then export CPTB200_LIB=$PWD/cuda-pt_b200/libcpt_b200$v.so; else unset CPTB200_LIB; fi
  echo "variant '$v'"; python tools/quick_bench.py scenes/xml/vader_wfpt_1080p.xml 32 2 16
done
