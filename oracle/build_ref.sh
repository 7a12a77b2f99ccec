#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Builds the reference itself as a checker.
#
# Compiles the reference's own headers (where they lie under /root/reference/src) together with
# oracle/ref_shim.cc into oracle/_ref/libccref.so (REF-FIXED: matrix.h:50 end() bug fixed, the
# oracle) and oracle/_ref/libccref_head.so (REF-HEAD: as shipped).  The reference targets
# clang+libc++; g++ 13 needs four mechanical portability patches.  They are
# applied to a scratch copy under $TMPDIR which is deleted afterwards: no reference source enters
# this repository, only the two built .so files land in oracle/_ref/ (git-ignored, gpurun-shipped).
#
# We do not run the reference's CMake build (clang/libc++ only, CMakeLists.txt:6,15-26).
# Flags: -O3 and NDEBUG *off* like the shipped configuration (CMakeLists.txt:7,11).
set -euo pipefail
REF=${CCREF_SRC:-/root/reference/src}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF" ]; then
  if [ -f "$OUT/libccref.so" ]; then
    echo "build_ref: $REF absent, keeping prebuilt $OUT/libccref.so"; exit 0
  fi
  echo "build_ref: $REF absent and no prebuilt library" >&2; exit 1
fi
mkdir -p "$OUT"
# up to date?
if [ -f "$OUT/libccref.so" ] && [ -f "$OUT/libccref_head.so" ] \
   && [ "$OUT/libccref.so" -nt "$HERE/ref_shim.cc" ] && [ "$OUT/libccref.so" -nt "$HERE/build_ref.sh" ] \
   && [ -f "$OUT/integration_test" ] && [ "$OUT/integration_test" -nt "$HERE/integration_test.cc" ] \
   && [ "$OUT/integration_test" -nt "$HERE/gpu_decoder.h" ] \
   && [ "${1:-}" != "--force" ]; then
  echo "build_ref: up to date"; exit 0
fi
SCRATCH=$(mktemp -d)
trap 'rm -rf "$SCRATCH"' EXIT
cp -r "$REF" "$SCRATCH/src"
chmod -R u+w "$SCRATCH/src"
python3 - "$SCRATCH/src" <<'EOF'
import re, sys, pathlib
src = pathlib.Path(sys.argv[1])
def patch(rel, fn):
    p = src / rel
    s = p.read_text()
    t = fn(s)
    assert t != s, f"patch for {rel} did not apply"
    p.write_text(t)
# 1. center.h: default template argument repeated on the out-of-class operator<< definition
def center(s):
    needle = "template <typename charT, typename traits = std::char_traits<charT> >\nstd::basic_ostream<charT, traits> &operator<<("
    return s.replace(needle, "template <typename charT, typename traits>\nstd::basic_ostream<charT, traits> &operator<<(")
patch("center.h", center)
# 2. galois.h: missing includes; std::array of an incomplete type
def galois(s):
    s = s.replace("#include <array>", "#include <array>\n#include <vector>\n#include <cstdint>\n#include <limits>\n#include <stdexcept>\n#include <utility>\n#include <ostream>", 1)
    s = s.replace("using Exp_table_type = std::array<ef_element, 2 * size>;", "using Exp_table_type = std::vector<ef_element>;")
    s = s.replace("typename GF::Exp_table_type exp;", "typename GF::Exp_table_type exp(2 * size);")
    return s
patch("math/galois.h", galois)
# 3. polynomial.h: make_tuple returned where a pair is declared
patch("math/polynomial.h", lambda s: s.replace("return std::make_tuple(", "return std::make_pair("))
EOF
CXXFLAGS="-std=c++17 -fpermissive -w -O3 -fPIC -pthread -ffp-contract=off"
SRCS="$HERE/ref_shim.cc $SCRATCH/src/codes/codes.c++ $SCRATCH/src/simulation/simulation.c++"
# REF-HEAD first (unfixed matrix.h)
g++ $CXXFLAGS -DCCREF_HEAD -I"$SCRATCH/src" -shared -o "$OUT/libccref_head.so" -x c++ $SRCS &
HEAD_PID=$!
# REF-FIXED: one-token fix of matrix<T>::end() const (matrix.h:50) in a second scratch copy
cp -r "$SCRATCH/src" "$SCRATCH/src_fixed"
python3 - "$SCRATCH/src_fixed/math/matrix.h" <<'EOF'
import sys
p = sys.argv[1]
s = open(p).read()
needle = "const_iterator end() const noexcept { return data.begin(); }"
assert needle in s
open(p, "w").write(s.replace(needle, "const_iterator end() const noexcept { return data.end(); }"))
EOF
SRCS_FIXED="$HERE/ref_shim.cc $SCRATCH/src_fixed/codes/codes.c++ $SCRATCH/src_fixed/simulation/simulation.c++"
g++ $CXXFLAGS -I"$SCRATCH/src_fixed" -shared -o "$OUT/libccref.so" -x c++ $SRCS_FIXED
wait $HEAD_PID
# the INTEGRATION.md adapter compiled against the real reference headers (needs libccgpu.so to link)
LIBDIR="$HERE/../channelcoding_b200"
if [ -f "$LIBDIR/libccgpu.so" ]; then
  g++ $CXXFLAGS -I"$SCRATCH/src_fixed" -I"$HERE" -I"$HERE/../include" "$HERE/integration_test.cc" \
      "$SCRATCH/src_fixed/codes/codes.c++" "$SCRATCH/src_fixed/simulation/simulation.c++" \
      -L"$LIBDIR" -lccgpu -Wl,-rpath,'$ORIGIN/../../channelcoding_b200' -o "$OUT/integration_test"
fi
ls -la "$OUT"
echo "build_ref: done"
