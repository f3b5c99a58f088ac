#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Builds the UNMODIFIED Omega2D reference hot path (USE_VC off, S=float, A=double) into
#   oracle/_ref/libo2dref_{rm,exp,wl,v2,v3}.so      (git-ignored, travels with gpurun)
# from the sources where they lie under $O2D_REFERENCE (default /root/reference).
#
# The core function is an unguarded "#define USE_*_KERNEL" inside src/CoreFunc.h
# (reference src/CoreFunc.h:35-39) and every header includes its siblings relative to
# its own directory, so per core the headers are staged in a throw-away directory
# under /tmp with that one line flipped by sed; the staging directory is deleted
# afterwards and nothing from the reference is ever written into this repository.
#
# Flags: the reference's Release flags are "-O3 -march=native" + OpenMP
# (reference CMakeLists.txt:55-57).  -march=native is replaced by -march=x86-64-v3
# (AVX2+FMA) because the objects are built in one container and executed on another
# host (the GPU box); asserts stay on, as in the reference's Release build.
set -euo pipefail
REF="${O2D_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
ROOT="$(cd "$HERE/.." && pwd)"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF/src not found; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$OUT"
MARCH="${O2D_REF_MARCH:--march=x86-64-v3}"
# -Bsymbolic / -fno-gnu-unique: the five per-core libraries export identically named template
# instantiations; bind each library's calls to its OWN copies even when several are loaded into one process
CXXFLAGS="-std=c++17 -O3 $MARCH -fopenmp -fPIC -DVERBOSE=false -w -fno-gnu-unique -Wl,-Bsymbolic"
STAGE="$(mktemp -d /tmp/o2dref.XXXXXX)"
trap 'rm -rf "$STAGE"' EXIT

# third-party helpers the container headers pull in transitively (oracle-dep)
gcc -O2 -fPIC -w -c "$REF/lib/tinyexpr/tinyexpr.c" -o "$STAGE/tinyexpr.o"
g++ -std=c++17 -O2 -fPIC -w -c "$REF/lib/tinyxml2/tinyxml2.cpp" -o "$STAGE/tinyxml2.o"

build_core () {
  local name="$1" macro="$2"
  local d="$STAGE/$name"
  mkdir -p "$d"
  cp "$REF"/src/*.h "$d"/
  cp -r "$REF"/src/shaders "$d"/
  cp "$REF"/src/Body.cpp "$d"/
  # exactly one core define active
  sed -i -E 's@^(#define USE_(RM|EXPONENTIAL|WL|V2|V3)_KERNEL)@//\1@' "$d/CoreFunc.h"
  sed -i -E "s@^//#define ${macro}\$@#define ${macro}@" "$d/CoreFunc.h"
  if [ "$(grep -c -E '^#define USE_(RM|EXPONENTIAL|WL|V2|V3)_KERNEL' "$d/CoreFunc.h")" != "1" ]; then
    echo "build_ref.sh: failed to select $macro" >&2; exit 1
  fi
  g++ $CXXFLAGS -I"$d" -I"$REF/lib" -I"$REF/lib/tinyexpr" -I"$REF/lib/tinyxml2" \
      -shared -o "$OUT/libo2dref_${name}.so" \
      "$HERE/ref_driver.cpp" "$d/Body.cpp" "$STAGE/tinyexpr.o" "$STAGE/tinyxml2.o"
  echo "built $OUT/libo2dref_${name}.so"
  # the drop-in check (tests/cpp/ref_dropin.cpp): the reference's OWN containers with this repository's influence headers
  # in place of Influence.h, linked against the product library, compared inside the program with the reference's CPU
  # loops from the library just built.  Default core only (V2), like an unmodified Omega2D build.
  if [ "$name" = "v2" ] && [ -f "$ROOT/omega2d_b200/libomega2d_b200.so" ]; then
    g++ -std=c++17 -O2 $MARCH -fopenmp -DVERBOSE=false -w -I"$d" -I"$REF/lib" -I"$REF/lib/tinyexpr" -I"$REF/lib/tinyxml2" \
        -I"$ROOT/omega2d_b200/host" "$ROOT/tests/cpp/ref_dropin.cpp" "$d/Body.cpp" "$STAGE/tinyexpr.o" "$STAGE/tinyxml2.o" \
        -L"$OUT" -lo2dref_v2 -L"$ROOT/omega2d_b200" -lomega2d_b200 \
        -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../omega2d_b200' -o "$OUT/ref_dropin_v2"
    echo "built $OUT/ref_dropin_v2"
    # the zero-source-change route: the reference's UNMODIFIED Influence.h with -DEXTERNAL_VEL_SOLVE, its two plugin
    # symbols resolved from the product library (tests/cpp/ref_external_seam.cpp)
    g++ -std=c++17 -O2 $MARCH -fopenmp -DVERBOSE=false -DEXTERNAL_VEL_SOLVE -w -I"$d" -I"$REF/lib" -I"$REF/lib/tinyexpr" -I"$REF/lib/tinyxml2" \
        "$ROOT/tests/cpp/ref_external_seam.cpp" "$d/Body.cpp" "$STAGE/tinyexpr.o" "$STAGE/tinyxml2.o" \
        -L"$OUT" -lo2dref_v2 -L"$ROOT/omega2d_b200" -lomega2d_b200 \
        -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../omega2d_b200' -o "$OUT/ref_external_seam_v2"
    echo "built $OUT/ref_external_seam_v2"
  fi
}

pids=()
build_core rm  USE_RM_KERNEL          & pids+=($!)
build_core exp USE_EXPONENTIAL_KERNEL & pids+=($!)
build_core wl  USE_WL_KERNEL          & pids+=($!)
build_core v2  USE_V2_KERNEL          & pids+=($!)
build_core v3  USE_V3_KERNEL          & pids+=($!)
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
exit $rc
