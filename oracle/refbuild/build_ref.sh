#!/usr/bin/env bash
# Builds the UNMODIFIED reference renderers into oracle/_ref/ as test oracles.
#
#   oracle/_ref/libref_rt.so                 raytracer  (resolution is run-time)
#   oracle/_ref/libref_rast_<W>x<H>.so       rasteriser (one per resolution: its
#                                            frame buffers are static arrays)
#   oracle/_ref/libprog_{rt,rast}_ref*.so    the reference PROGRAMS (main + Update + Draw) headless
#   oracle/_ref/libprog_{rt,rast}_dropin*.so the same programs with this repository's Draw shim in
#                                            place of the reference's Draw, linked to libb200render.so
#                                            (prog_harness.cpp; the executed drop-in test)
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

# ---- the programs, with the reference's Draw and with the drop-in Draw -----------
# Only the DEFINITION line of Draw is renamed for the drop-in flavour (RT skeleton.cpp:104, RAST
# :203); its declaration and the call in main() stay, and now resolve to the shim's Draw.
PKG="$HERE/../../computer-graphics_b200"
PROG_SIZES="${PROG_SIZES:-900x720 64x48}"
RENAME='s/^void Draw(screen\* screen)\([^;]*\)$/void reference_Draw(screen* screen)\1/'
PFLAGS="-O3 -pipe -w -fPIC -shared -std=c++11 -DB200_SDL_SCRIPTED"
# The programs print through iostream: they must share the process's libstdc++ (a compiler wrapper
# that links it statically gives the library a second, half-preempted copy inside Python).
PCXX="${PROG_CXX:-$( [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo "$CXX" )}"
LINK="-L$PKG -lb200render -Wl,-rpath,\$ORIGIN/../../computer-graphics_b200"
mkdir -p "$TMP/rt_ref" "$TMP/rt_dropin" "$TMP/rast_ref" "$TMP/rast_dropin"
cp "$TMP/skeleton_rt_patched.cpp" "$TMP/rt_ref/skeleton_prog.cpp"
sed -e "$RENAME" "$TMP/skeleton_rt_patched.cpp" > "$TMP/rt_dropin/skeleton_prog.cpp"
cp "$TMP/skeleton_rast_patched.cpp" "$TMP/rast_ref/skeleton_prog.cpp"
sed -e "$RENAME" "$TMP/skeleton_rast_patched.cpp" > "$TMP/rast_dropin/skeleton_prog.cpp"
grep -q "^void reference_Draw" "$TMP/rt_dropin/skeleton_prog.cpp" && grep -q "^void reference_Draw" "$TMP/rast_dropin/skeleton_prog.cpp" \
  || { echo "build_ref: could not rename the reference's Draw definition" >&2; exit 1; }
RT_INC="-I$HERE/stubs -I$REF/raytracer/Source -I$REF/glm"
RAST_INC="-I$HERE/stubs -I$REF/rasteriser/Source -I$REF/glm"
$PCXX $PFLAGS -DPROG_RT -I"$TMP/rt_ref" $RT_INC "$HERE/prog_harness.cpp" -o "$OUT/libprog_rt_ref.so" &
for sz in $PROG_SIZES; do
  W="${sz%x*}"; H="${sz#*x}"
  $PCXX $PFLAGS -mcmodel=medium -DPROG_RAST -DREF_W="$W" -DREF_H="$H" -I"$TMP/rast_ref" $RAST_INC \
      "$HERE/prog_harness.cpp" -o "$OUT/libprog_rast_ref_${W}x${H}.so" &
done
if [ -f "$PKG/libb200render.so" ]; then
  SHIM="-I$PKG/host/shim -I$HERE/../../include"
  $PCXX $PFLAGS -DPROG_RT -DDROPIN -I"$TMP/rt_dropin" $RT_INC $SHIM "$HERE/prog_harness.cpp" $LINK -o "$OUT/libprog_rt_dropin.so" &
  for sz in $PROG_SIZES; do
    W="${sz%x*}"; H="${sz#*x}"
    $PCXX $PFLAGS -mcmodel=medium -DPROG_RAST -DDROPIN -DREF_W="$W" -DREF_H="$H" -I"$TMP/rast_dropin" $RAST_INC $SHIM \
        "$HERE/prog_harness.cpp" $LINK -o "$OUT/libprog_rast_dropin_${W}x${H}.so" &
  done
else
  echo "build_ref: $PKG/libb200render.so not built yet; drop-in programs skipped" >&2
fi
wait
ls "$OUT"/libprog_* | sed 's/^/built /'
