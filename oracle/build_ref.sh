#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the *unmodified* reference (MatthewBonanni/Mallard) CPU code
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored).
#
#  * Third-party, vendored deps (Kokkos 4.5.1, KokkosKernels 4.5.1 BLAS-1 only) are configured
#    with their own cmake (they need a generated config header); Serial + OpenMP back-ends.
#  * Mallard's own sources are compiled DIRECTLY with g++ below (its top-level CMakeLists needs
#    HDF5 + a network-fetched gtest and mis-spells the FP64 macro, so it is not used).
#  * -DMallard_USE_DOUBLE: the reference's CMake defines Mallard_USE_DOUBLES (typo) and thus
#    builds FP32; the path we replace is FP64 (src/common/common_typedef.h:27-31).
#
# Outputs (only under oracle/_ref/):  lib/libmallard_ref.a, bin/Mallard (stock driver),
# bin/ref_harness (oracle/ref_harness.cpp: stage-level dumps + timing), bin/MallardTest.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${MALLARD_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
BLD="${MALLARD_ORACLE_BUILD:-/tmp/mallard_oracle_build}"
CXX_BIN=/usr/bin/g++            # $CXX=/opt/gcc/bin/g++ lacks libgomp.spec
JOBS="${JOBS:-$(nproc)}"
[ -d "$REF/src" ] || { echo "reference not present at $REF; keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT" "$BLD"
P="$OUT/kokkos"

if [ ! -f "$P/lib/libkokkoscore.a" ]; then
  cmake -S "$REF/src/external/kokkos" -B "$BLD/kk" -G Ninja -DCMAKE_CXX_COMPILER=$CXX_BIN \
    -DCMAKE_BUILD_TYPE=Release -DCMAKE_CXX_STANDARD=20 -DKokkos_ENABLE_SERIAL=ON \
    -DKokkos_ENABLE_OPENMP=ON -DBUILD_SHARED_LIBS=OFF -DCMAKE_INSTALL_LIBDIR=lib \
    -DCMAKE_INSTALL_PREFIX="$P" >/dev/null
  ninja -C "$BLD/kk" -j"$JOBS" install >/dev/null
fi
if [ ! -f "$P/lib/libkokkoskernels.a" ]; then
  cmake -S "$REF/src/external/kokkos-kernels" -B "$BLD/kkk" -G Ninja -DCMAKE_CXX_COMPILER=$CXX_BIN \
    -DCMAKE_BUILD_TYPE=Release -DKokkos_DIR="$P/lib/cmake/Kokkos" -DCMAKE_INSTALL_PREFIX="$P" \
    -DBUILD_SHARED_LIBS=OFF -DCMAKE_INSTALL_LIBDIR=lib \
    -DKokkosKernels_ENABLE_ALL_COMPONENTS=OFF -DKokkosKernels_ENABLE_COMPONENT_BLAS=ON \
    -DKokkosKernels_INST_DOUBLE=ON >/dev/null
  ninja -C "$BLD/kkk" -j"$JOBS" install >/dev/null
fi

INC="-I$P/include"
for d in solver numerics mesh common boundary physics io; do INC="$INC -I$REF/src/$d"; done
INC="$INC -I$REF/src/external/toml11/include -isystem $REF/src/external/exprtk"
CXXFLAGS="-std=c++20 -O3 -fopenmp -DMallard_USE_DOUBLE -DNDEBUG -w"
LIBS="-L$P/lib -lkokkoskernels -lkokkoscontainers -lkokkoscore -lkokkossimd -ldl"

mkdir -p "$BLD/obj" "$OUT/lib" "$OUT/bin"
OBJS=()
pids=()
n=0
for f in $(cd "$REF/src" && ls solver/*.cpp numerics/*.cpp mesh/*.cpp common/*.cpp boundary/*.cpp physics/*.cpp io/*.cpp); do
  o="$BLD/obj/$(echo "$f" | tr '/' '_').o"
  OBJS+=("$o")
  if [ ! -f "$o" ] || [ "$REF/src/$f" -nt "$o" ]; then
    $CXX_BIN $CXXFLAGS $INC -c "$REF/src/$f" -o "$o" &
    pids+=($!); n=$((n+1))
    if [ $n -ge "$JOBS" ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); n=$((n-1)); fi
  fi
done
wait
rm -f "$OUT/lib/libmallard_ref.a"
ar rcs "$OUT/lib/libmallard_ref.a" "${OBJS[@]}"

$CXX_BIN $CXXFLAGS $INC "$REF/src/main.cpp" "$OUT/lib/libmallard_ref.a" $LIBS -o "$OUT/bin/Mallard"
if [ -f "$HERE/ref_harness.cpp" ]; then
  $CXX_BIN $CXXFLAGS $INC "$HERE/ref_harness.cpp" "$OUT/lib/libmallard_ref.a" $LIBS -o "$OUT/bin/ref_harness"
fi
# the reference host driving the B200 library through its C ABI (integration proof)
if [ -f "$HERE/dropin_harness.cpp" ] && [ -f "$HERE/../mallard_b200/libmallard_b200.so" ]; then
  $CXX_BIN $CXXFLAGS $INC "$HERE/dropin_harness.cpp" "$OUT/lib/libmallard_ref.a" $LIBS \
     -L"$HERE/../mallard_b200" -lmallard_b200 -Wl,-rpath,'$ORIGIN/../../../mallard_b200' -o "$OUT/bin/mallard_dropin"
fi
if [ "${BUILD_REF_TESTS:-0}" = 1 ]; then
  G="$REF/src/external/kokkos/tpls/gtest"
  $CXX_BIN $CXXFLAGS $INC -I"$G" -I"$REF/test" "$REF"/test/*.cpp "$G/gtest/gtest-all.cc" \
     "$OUT/lib/libmallard_ref.a" $LIBS -lpthread -o "$OUT/bin/MallardTest"
fi
echo "oracle/_ref built: $(ls "$OUT/bin")"
