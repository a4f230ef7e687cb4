#!/bin/bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference sources where they lie under /root/reference (nothing is
# copied) against the MKL shim + the image's OpenBLAS 0.3.15, into oracle/_ref/ (git-ignored,
# travels to the GPU box).  Also builds the standalone RNG helper used by the numpy twin.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${RSVD_REFERENCE_DIR:-/root/reference}"
PYLIBS="$(python -c 'import sys,glob,os; import site; print([p for p in site.getsitepackages() if os.path.isdir(os.path.join(p,"opencv_python_headless.libs"))][0])')/opencv_python_headless.libs"
OPENBLAS="$(ls "$PYLIBS"/libopenblasp-r0-*.so | head -1)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
CF="-O2 -fopenmp -w -fPIC -shared -ffp-contract=off -I$HERE/shim -include $HERE/shim/zero_malloc.h"
LF="$OPENBLAS -Wl,-rpath,$PYLIBS -Wl,-rpath-link,$PYLIBS -Wl,-Bsymbolic -Wl,--disable-new-dtags -lm"
# RNG helper (no reference sources involved)
gcc -O2 -fopenmp -fPIC -shared -ffp-contract=off -I$HERE/shim "$HERE/shim/vsl_shim.c" -lm -o "$OUT/liboracle_rng.so"
if [ -d "$REF/multi_core_mkl_code" ]; then
  D32="$REF/multi_core_mkl_code"; D64="$REF/multi_core_mkl_code_64bit"
  gcc $CF -I"$D32" "$HERE/shim/vsl_shim.c" "$D32/rank_revealing_algorithms_intel_mkl.c" \
      "$D32/matrix_vector_functions_intel_mkl.c" $LF -o "$OUT/libref32.so"
  gcc $CF -I"$D64" "$HERE/shim/vsl_shim.c" "$D64/rank_revealing_algorithms_intel_mkl.c" \
      "$D64/matrix_vector_functions_intel_mkl.c" $LF -o "$OUT/libref64.so"
  echo "built $OUT/libref32.so $OUT/libref64.so against $OPENBLAS"
else
  echo "reference tree $REF absent: keeping prebuilt oracle/_ref/*.so"
fi
