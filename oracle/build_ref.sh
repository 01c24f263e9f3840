#!/usr/bin/env bash
# Compiles the UNMODIFIED reference (facebookresearch/dietgpu) CUDA + host
# sources, from where they lie under $REF (default /root/reference), for
# sm_100a into oracle/_ref/libdietgpu_ref.so, together with oracle/ref_harness.cu
# (a C-ABI veneer written for this repo) and the glog stand-in in
# oracle/ref_shim/.  Drives nvcc directly -- the reference's own CMake build
# is NOT run.  Outputs only into oracle/_ref/ (git-ignored, travels to the GPU
# box with gpurun).  TEST / BASELINE infrastructure only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/dietgpu" ]; then
  echo "build_ref: $REF/dietgpu not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 -O2 -DNDEBUG -gencode arch=compute_100a,code=sm_100a -lineinfo
       -I"$REF" -I"$HERE/ref_shim" -Xcompiler -fPIC -ccbin /usr/bin/g++ -w)
SRCS=(dietgpu/ans/GpuANSEncode.cu dietgpu/ans/GpuANSDecode.cu dietgpu/ans/GpuANSInfo.cu
      dietgpu/float/GpuFloatCompress.cu dietgpu/float/GpuFloatDecompress.cu dietgpu/float/GpuFloatInfo.cu)
pids=()
for s in "${SRCS[@]}"; do
  o="$OUT/obj/$(basename "${s%.cu}").o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$REF/$s" -o "$o" &
    pids+=($!)
  fi
done
for s in dietgpu/utils/DeviceUtils.cpp dietgpu/utils/StackDeviceMemory.cpp; do
  o="$OUT/obj/$(basename "${s%.cpp}").o"
  if [ ! -f "$o" ]; then
    "$NVCC" "${FLAGS[@]}" -x cu -c "$REF/$s" -o "$o" &
    pids+=($!)
  fi
done
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_harness.cu" -o "$OUT/obj/ref_harness.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$OUT/libdietgpu_ref.so" "$OUT"/obj/*.o -lcudart -ccbin /usr/bin/g++
echo "built $OUT/libdietgpu_ref.so"
