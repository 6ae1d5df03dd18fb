#!/usr/bin/env bash
# Compile-checks the two Draw shims INSIDE the reference's own translation units
# (needs /root/reference; run in the dev container).  The reference's Draw is
# renamed out of the way by macro, then the shim is included, exactly as
# INTEGRATION.md describes.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../../.."
REF="${REF:-/root/reference}"
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
for which in rt rast; do
  dir=$([ $which = rt ] && echo raytracer || echo rasteriser)
  sed -e 's|^#include "/usr/local/Cellar/opencv[^"]*"|#include "cv_stub.hpp"|' "$REF/$dir/Source/skeleton.cpp" > "$TMP/skeleton.cpp"
  cat > "$TMP/tu_$which.cpp" <<TU
#define Draw reference_Draw
#include "skeleton.cpp"
#undef Draw
#include "draw_$which.inc"
TU
  g++ -std=c++11 -w -fsyntax-only -I"$ROOT/oracle/refbuild/stubs" -I"$TMP" -I"$HERE" -I"$REF/$dir/Source" \
      -I"$REF/glm" -I"$ROOT/include" "$TMP/tu_$which.cpp"
  echo "shim $which: compiles inside $dir/Source/skeleton.cpp"
done
