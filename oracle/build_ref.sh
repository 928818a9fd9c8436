#!/usr/bin/env bash
# Build the UNMODIFIED reference (Enigmatisms/cuda-pt) from the sources where
# they lie under /root/reference, plus our thin driver oracle/ref_harness/ref_dump.cu,
# into oracle/_ref/ref_dump (git-ignored; travels to the GPU box with gpurun).
#
# The reference's own CMake build is NOT used (it hard-requires OpenGL/GLEW dev
# packages which this image lacks, CMakeLists.txt:90-97); this script compiles
# the same translation units with the same CUDA flags (CMakeLists.txt:7,52,66,72)
# at -arch=sm_100a.  The only shim is a 2-line GL/gl.h that forwards to the
# reference's vendored glad header, needed because renderer/path_tracer.cuh:30
# includes <cuda_gl_interop.h>.
#
# Test infrastructure only: nothing here is linked into the product library.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
OUT="$HERE/_ref"
OBJ="$OUT/obj"
if [ ! -d "$REF/src" ]; then
  echo "[build_ref] $REF not present: keeping prebuilt oracle/_ref (if any)"; exit 0
fi
mkdir -p "$OBJ" "$OUT/glshim/GL"
printf '#pragma once\n#include <glad/gl.h>\n' > "$OUT/glshim/GL/gl.h"

INC="-I$OUT/glshim -I$REF/src/ext/glfw/deps -I$REF/src -I$REF/src/ext/tinyxml2 \
 -I$REF/src/ext/tinyobjloader -I$REF/src/ext -I$REF/src/ext/atomic_queue/include -I$ROOT/include"
NVFLAGS="-std=c++20 --expt-relaxed-constexpr --extended-lambda --use_fast_math -Xptxas -O3 -O3 \
 --maxrregcount=40 -lineinfo -gencode arch=compute_100a,code=sm_100a -rdc=true \
 -Xcompiler -fopenmp,-fPIC -DOPENMP_ENABLED -w"

MK="$OBJ/Makefile"
{
  echo "all: $OUT/ref_dump"
  OBJS=""
  for f in "$REF"/src/impl/*.cu "$REF"/src/impl/*.cpp "$REF"/src/pt_impl/*.cu "$HERE/ref_harness/ref_dump.cu"; do
    b=$(basename "$f"); o="$OBJ/${b%.*}.o"; OBJS="$OBJS $o"
    echo "$o: $f"
    printf '\tnvcc %s %s -x cu -c %s -o %s\n' "$NVFLAGS" "$INC" "$f" "$o"
  done
  echo "$OBJ/tinyxml2.o: $REF/src/ext/tinyxml2/tinyxml2.cpp"
  printf '\tg++ -O2 -fPIC -c $< -o $@\n'
  echo "$OBJ/tiny_obj_loader.o: $REF/src/ext/tinyobjloader/tiny_obj_loader.cc"
  printf '\tg++ -O2 -fPIC -x c++ -c $< -o $@\n'
  OBJS="$OBJS $OBJ/tinyxml2.o $OBJ/tiny_obj_loader.o"
  echo "$OUT/ref_dump: $OBJS"
  printf '\tnvcc -gencode arch=compute_100a,code=sm_100a -rdc=true %s -o $@ -lgomp\n' "$OBJS"
} > "$MK"
make -f "$MK" -j"$(nproc)" all
echo "[build_ref] built $OUT/ref_dump"
