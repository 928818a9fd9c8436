#!/usr/bin/env bash
# Build the REFERENCE'S OWN entry points - the nanobind module `pyrender` (src/pyrender/*.cu) and the offline renderer `pt`
# (app/pt_renderer.cu) - with their path-tracer switch cases routed to libcpt_b200.so through the adapter
# integration/b200_wavefront.cuh.  This is the change INTEGRATION.md describes, done mechanically:
#
#   * nothing under $REF (default /root/reference) is modified and no reference source is copied into this repository: the
#     two translation units that hold the switch cases are patched with `sed` INTO oracle/_ref/b200_side/ at build time
#     (git-ignored, like everything under oracle/_ref/), every other reference TU is the object oracle/build_ref.sh built;
#   * the patch = one #include + the constructor expressions of the MegaKernelPT / WavefrontPT / MegaKernelVPT cases
#     (python_render.cu:94-119, pt_renderer.cu:48-75); loader, SBVH builder, texture loading, light tracer, visualisers
#     stay the reference's.
#
# Outputs (travel to the GPU box with gpurun):
#   oracle/_ref/b200_side/pyrender.so   import as `pyrender` with this directory on sys.path
#   oracle/_ref/b200_side/pt            ./pt <scene.xml> -> render.png
#   oracle/_ref/pyrender_ref/pyrender.so  the UNPATCHED reference module (same objects, its own switch cases): the baseline of
#                                       the "Python serial path" timing in bench.py
# Test infrastructure for the boundary (tests/test_gpu.py::test_reference_pyrender_*): it proves the adapter compiles and runs
# inside the reference's build; it is not part of the product library.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
OUT="$ROOT/oracle/_ref/b200_side"
OBJ="$ROOT/oracle/_ref/obj"
if [ ! -d "$REF/src" ]; then
  echo "[reference-side] $REF not present: keeping prebuilt oracle/_ref/b200_side (if any)"; exit 0
fi
if [ ! -f "$OBJ/scene.o" ]; then
  bash "$ROOT/oracle/build_ref.sh"
fi
mkdir -p "$OUT/obj"
PYINC=$(python3 -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
NB="$REF/src/ext/nanobind"
INC="-I$ROOT/oracle/_ref/glshim -I$REF/src/ext/glfw/deps -I$REF/src -I$REF/src/ext/tinyxml2 \
 -I$REF/src/ext/tinyobjloader -I$REF/src/ext -I$REF/src/ext/atomic_queue/include -I$ROOT/include -I$HERE \
 -I$NB/include -I$NB/ext/robin_map/include -I$PYINC -I$REF/src/pyrender"
NVFLAGS="-std=c++20 --expt-relaxed-constexpr --extended-lambda --use_fast_math -Xptxas -O3 -O3 \
 --maxrregcount=40 -lineinfo -gencode arch=compute_100a,code=sm_100a -rdc=true \
 -Xcompiler -fopenmp,-fPIC -DOPENMP_ENABLED -w"

# ---- the patched translation units (generated, never committed)
patch_cases() { # $1 = source, $2 = destination, $3 = scene expression, $4 = xml path expression, $5 = device expression
  sed -e "0,/#include/s|#include|#include \"b200_wavefront.cuh\"\n#include|" \
      -e "s|std::make_unique<PathTracer<SingleTileScheduler>>($3)|std::make_unique<B200WavefrontTracer>($3, $4, $5, 1)|" \
      -e "s|std::make_unique<WavefrontPathTracer>($3)|std::make_unique<B200WavefrontTracer>($3, $4, $5, 0)|" \
      -e "s|std::make_unique<VolumePathTracer>($3)|std::make_unique<B200WavefrontTracer>($3, $4, $5, 2)|" "$1" > "$2"
  grep -c "B200WavefrontTracer" "$2" | grep -q "^3$" || { echo "[reference-side] patch of $1 did not apply (reference changed?)"; exit 1; }
}
patch_cases "$REF/src/pyrender/python_render.cu" "$OUT/python_render_b200.cu" '\*scene' 'path' '_device_id'
patch_cases "$REF/app/pt_renderer.cu" "$OUT/pt_renderer_b200.cu" 'scene' 'xml_path' '0'

CORE_OBJS=$(ls "$OBJ"/*.o | grep -v ref_dump.o)
MK="$OUT/obj/Makefile"
{
  echo "all: $OUT/pyrender.so $OUT/pt $ROOT/oracle/_ref/pyrender_ref/pyrender.so"
  echo "$OUT/obj/python_render_ref.o: $REF/src/pyrender/python_render.cu"
  printf '\tnvcc %s %s -x cu -c $< -o $@\n' "$NVFLAGS" "$INC"
  PYOBJS=""
  for f in "$OUT/python_render_b200.cu" "$REF/src/pyrender/python_bind.cu"; do
    b=$(basename "$f"); o="$OUT/obj/${b%.*}.o"; PYOBJS="$PYOBJS $o"
    echo "$o: $f $HERE/b200_wavefront.cuh $HERE/cpts_flatten.cuh"
    printf '\tnvcc %s %s -x cu -c %s -o %s\n' "$NVFLAGS" "$INC" "$f" "$o"
  done
  for f in nb_func common error nb_internals nb_type nb_static_property implicit nb_ndarray; do
    o="$OUT/obj/$f.o"; PYOBJS="$PYOBJS $o"
    echo "$o: $NB/src/$f.cpp"
    printf '\tg++ -std=c++17 -O2 -fPIC -fvisibility=hidden -DNDEBUG -I%s/include -I%s/ext/robin_map/include -I%s -c $< -o $@\n' "$NB" "$NB" "$PYINC"
  done
  echo "$OUT/obj/pt_renderer_b200.o: $OUT/pt_renderer_b200.cu $HERE/b200_wavefront.cuh $HERE/cpts_flatten.cuh"
  printf '\tnvcc %s %s -x cu -c $< -o $@\n' "$NVFLAGS" "$INC"
  echo "$OUT/pyrender.so: $PYOBJS"
  printf '\tnvcc -gencode arch=compute_100a,code=sm_100a -rdc=true -shared -Xcompiler -fPIC %s %s -o $@ -L%s -lcpt_b200 -lgomp -Xlinker -rpath=%s\n' \
         "$PYOBJS" "$(echo $CORE_OBJS)" "$ROOT/cuda-pt_b200" "\\\$\$ORIGIN/../../../cuda-pt_b200"
  NBOBJS=$(for f in nb_func common error nb_internals nb_type nb_static_property implicit nb_ndarray; do printf '%s ' "$OUT/obj/$f.o"; done)
  echo "$ROOT/oracle/_ref/pyrender_ref/pyrender.so: $OUT/obj/python_render_ref.o $OUT/obj/python_bind.o $NBOBJS"
  printf '\tmkdir -p %s && nvcc -gencode arch=compute_100a,code=sm_100a -rdc=true -shared -Xcompiler -fPIC $^ %s -o $@ -lgomp\n' \
         "$ROOT/oracle/_ref/pyrender_ref" "$(echo $CORE_OBJS)"
  echo "$OUT/pt: $OUT/obj/pt_renderer_b200.o"
  printf '\tnvcc -gencode arch=compute_100a,code=sm_100a -rdc=true $< %s -o $@ -L%s -lcpt_b200 -lgomp -Xlinker -rpath=%s\n' \
         "$(echo $CORE_OBJS)" "$ROOT/cuda-pt_b200" "\\\$\$ORIGIN/../../../cuda-pt_b200"
} > "$MK"
make -f "$MK" -j"$(nproc)" all
echo "[reference-side] built $OUT/pyrender.so and $OUT/pt"
