#!/bin/bash
# Builds the reference's own gtests, UNCHANGED and where they lie under /root/reference, against the
# product (libdietgpu_b200.so) through tests/ref_shim -> include/dietgpu_b200_compat.hpp.
# Outputs (git-ignored binaries that travel to the GPU box): tests/cpp/ref_ans_test, ref_float_test.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF=/root/reference
GT=$REF/third_party/googletest/googletest
[ -d "$REF/dietgpu" ] || { echo "no /root/reference here: keeping prebuilt binaries"; exit 0; }
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OBJ="$HERE/obj"
mkdir -p "$OBJ"
if [ ! -f "$OBJ/gtest-all.o" ]; then
  g++ -std=c++17 -O1 -I"$GT/include" -I"$GT" -c "$GT/src/gtest-all.cc" -o "$OBJ/gtest-all.o"
  g++ -std=c++17 -O1 -I"$GT/include" -I"$GT" -c "$GT/src/gtest_main.cc" -o "$OBJ/gtest_main.o"
fi
FLAGS="-std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -I$ROOT/tests/ref_shim -I$ROOT/include -I$GT/include"
LINK="-L$ROOT/dietgpu_b200 -ldietgpu_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../dietgpu_b200 -lcudart -lpthread"
$NVCC $FLAGS -o "$HERE/ref_ans_test" "$REF/dietgpu/ans/ANSTest.cu" "$OBJ/gtest-all.o" "$OBJ/gtest_main.o" $LINK
$NVCC $FLAGS -o "$HERE/ref_float_test" "$REF/dietgpu/float/FloatTest.cu" "$OBJ/gtest-all.o" "$OBJ/gtest_main.o" $LINK
echo "built $HERE/ref_ans_test $HERE/ref_float_test"
