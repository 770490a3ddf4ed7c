#!/bin/bash
# oracle/ref_build.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the pieces of the reference that build from their own few source files (no cmake, no external libraries)
# into oracle/_ref/libtetwild_ref.so, reading the sources where they lie under $REF (default /root/reference).
# Nothing from the reference is copied into tracked files: line-range extracts go to oracle/_ref/gen/ (git-ignored).
# The whole TetWild binary is NOT buildable here (needs cmake-downloaded geogram/libigl/CGAL/Boost/GMP): see DESIGN.md.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
CC=$( [ -x /usr/bin/gcc ] && echo /usr/bin/gcc || echo gcc )
CXX=$( [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++ )
if [ ! -d "$REF/src/tetwild" ]; then
  echo "ref_build: $REF not present; keeping any prebuilt $OUT/libtetwild_ref.so" >&2
  exit 0
fi
mkdir -p "$OUT/gen"
# (1) AMIPS energy/Jacobian/Hessian: LocalOperations.cpp:28-291
sed -n '28,291p' "$REF/src/tetwild/LocalOperations.cpp" > "$OUT/gen/amips_lines.inc"
# (2) sampleTriangle: Common.cpp:143 up to the closing brace before '} // namespace tetwild'
awk 'NR>=143 && /^} \/\/ namespace tetwild/ {exit} NR>=143 {print}' "$REF/src/tetwild/Common.cpp" > "$OUT/gen/sample_lines.inc"
CXXFLAGS="-O2 -fPIC -fopenmp -ffp-contract=off -std=c++14 -w"
$CXX $CXXFLAGS -I"$HERE/shim" -I"$REF/src" -I"$OUT/gen" -c "$REF/src/tetwild/geogram/mesh_AABB.cpp" -o "$OUT/gen/mesh_AABB.o"
$CXX $CXXFLAGS -I"$HERE/shim" -I"$REF/src" -I"$OUT/gen" -c "$HERE/ref_wrap.cpp" -o "$OUT/gen/ref_wrap.o"
$CC -O2 -fPIC -fopenmp -ffp-contract=off -std=gnu11 -I"$HERE" -c "$HERE/envelope.c" -o "$OUT/gen/envelope_leaf.o"
$CC -O2 -fPIC -fopenmp -ffp-contract=off -std=gnu11 -I"$HERE" -c "$HERE/predicates.c" -o "$OUT/gen/predicates.o"
$CXX -shared -fopenmp -o "$OUT/libtetwild_ref.so" "$OUT/gen/mesh_AABB.o" "$OUT/gen/ref_wrap.o" "$OUT/gen/envelope_leaf.o" "$OUT/gen/predicates.o" -lm
echo "ref_build: built $OUT/libtetwild_ref.so"
# (4) the same AMIPS text in IEEE binary128 (higher-precision truth for the 1e-9 question): oracle/ref_quad.cpp
$CXX -O2 -fPIC -fopenmp -std=c++14 -w -I"$OUT/gen" -shared -o "$OUT/libtetwild_ref_quad.so" "$HERE/ref_quad.cpp" -lquadmath
echo "ref_build: built $OUT/libtetwild_ref_quad.so"
