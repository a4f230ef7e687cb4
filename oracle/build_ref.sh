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
  # The reference's drivers 2 and 4 linked with the reference itself (libref32.so): the other arm of the driver-level parity
  # test, whose B200 arm is the relinked binary below.
  mkdir -p "$OUT/refdrv"
  for d in driver_multi_core_mkl2 driver_multi_core_mkl4; do
    gcc -O2 -fopenmp -w -ffp-contract=off -I"$HERE/shim" -include "$HERE/shim/zero_malloc.h" -I"$D32" "$D32/$d.c" -L"$OUT" -lref32 \
        -Wl,-rpath,'$ORIGIN/..' -Wl,-rpath,$PYLIBS -Wl,--disable-new-dtags -lm -o "$OUT/refdrv/$d"
  done
  # Relink test: the reference's own drivers, UNMODIFIED, compiled against the drop-in headers (include/) and linked
  # with the B200 libraries instead of MKL.  The sources are compiled from a scratch copy (a quoted #include searches the
  # including file's directory first, where the reference's own headers live); only the binaries land in oracle/_ref/relink.
  PKG="$HERE/../lowrankmatrixdecompositioncodes_b200"
  if [ -f "$PKG/librsvd_b200_api32.so" ]; then
    mkdir -p "$OUT/relink"; TMP="$(mktemp -d)"
    for d in driver_multi_core_mkl1 driver_multi_core_mkl2 driver_multi_core_mkl3 driver_multi_core_mkl4 driver_multi_core_mkl5; do
      cp "$D32/$d.c" "$TMP/$d.c"
      gcc -O2 -w -I"$HERE/../include" "$TMP/$d.c" -L"$PKG" -lrsvd_b200_api32 -lrsvd_b200 -lm -Wl,-rpath,'$ORIGIN/../../../lowrankmatrixdecompositioncodes_b200' -o "$OUT/relink/$d"
    done
    cp "$D64/driver1.c" "$TMP/driver1_64bit.c"
    gcc -O2 -w -I"$HERE/../include/64bit" -I"$HERE/../include" "$TMP/driver1_64bit.c" -L"$PKG" -lrsvd_b200_api64 -lrsvd_b200 -lm -Wl,-rpath,'$ORIGIN/../../../lowrankmatrixdecompositioncodes_b200' -o "$OUT/relink/driver1_64bit"
    rm -rf "$TMP"
    echo "relinked reference drivers: $(ls $OUT/relink | tr '\n' ' ')"
  fi
else
  echo "reference tree $REF absent: keeping prebuilt oracle/_ref/*.so"
fi
