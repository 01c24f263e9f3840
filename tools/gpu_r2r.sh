#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export DIETGPU_B200_LIB=$PWD/dietgpu_b200/libdietgpu_b200_dw.so
for wl in c3 c2 c4; do
  WALL_GRAPH=0 timeout 300 python tools/walltime.py $wl "" "decode_warps=10" "decode_warps=20" "" 2>&1 | grep -v Warning
done | tee gpurun_out/r2r_wall.txt
