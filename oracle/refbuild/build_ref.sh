#!/usr/bin/env bash
# Builds the UNMODIFIED reference renderers into oracle/_ref/ as test oracles.
#
#   oracle/_ref/libref_rt.so                 raytracer  (resolution is run-time)
#   oracle/_ref/libref_rast_<W>x<H>.so       rasteriser (one per resolution: its
#                                            frame buffers are static arrays)
#
# The reference sources are compiled where they lie under $REF (default
# /root/reference); nothing is copied into the repository.  The only edits are
# made on the fly by sed into a temporary directory:
#   * the SCREEN_WIDTH / SCREEN_HEIGHT #defines (RT skeleton.cpp:19-20, RAST
#     skeleton.cpp:21-22) -- "-D" cannot override an in-file #define
#   * RAST skeleton.cpp:9, an absolute macOS OpenCV include -> inert stub
# Flags are the reference's own (raytracer/Makefile:15: -O3, no -march => no FMA
# contraction on x86-64), plus -fPIC -shared.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
REF="${REF:-/root/reference}"
RAST_SIZES="${RAST_SIZES:-900x720 3840x2160 320x240 64x48}"

if [ ! -d "$REF/raytracer/Source" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

CXX="${CXX:-g++}"
FLAGS="-O3 -pipe -w -fPIC -shared -std=c++11"

# ---- raytracer ---------------------------------------------------------------
sed -e 's/^#define SCREEN_WIDTH .*/#define SCREEN_WIDTH ref_screen_w/' \
    -e 's/^#define SCREEN_HEIGHT .*/#define SCREEN_HEIGHT ref_screen_h/' \
    "$REF/raytracer/Source/skeleton.cpp" > "$TMP/skeleton_rt_patched.cpp"
$CXX $FLAGS -I"$HERE/stubs" -I"$TMP" -I"$REF/raytracer/Source" -I"$REF/glm" \
    "$HERE/ref_rt_harness.cpp" -o "$OUT/libref_rt.so"
echo "built $OUT/libref_rt.so"

# ---- rasteriser --------------------------------------------------------------
sed -e 's/^#define SCREEN_WIDTH .*/#define SCREEN_WIDTH REF_W/' \
    -e 's/^#define SCREEN_HEIGHT .*/#define SCREEN_HEIGHT REF_H/' \
    -e 's|^#include "/usr/local/Cellar/opencv[^"]*"|#include "cv_stub.hpp"|' \
    "$REF/rasteriser/Source/skeleton.cpp" > "$TMP/skeleton_rast_patched.cpp"
for sz in $RAST_SIZES; do
  W="${sz%x*}"; H="${sz#*x}"
  $CXX $FLAGS -mcmodel=medium -DREF_W="$W" -DREF_H="$H" \
      -I"$HERE/stubs" -I"$TMP" -I"$REF/rasteriser/Source" -I"$REF/glm" \
      "$HERE/ref_rast_harness.cpp" -o "$OUT/libref_rast_${W}x${H}.so" &
done
wait
for sz in $RAST_SIZES; do echo "built $OUT/libref_rast_${sz}.so"; done
