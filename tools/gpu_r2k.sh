#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for wl in c3 c4 c2 c3x1; do
timeout 300 python tools/walltime.py $wl "" "decode_warps=4" "encode_wide_table=0" "encode_wide_table=1" 2>&1 | tee gpurun_out/r2k_wall_$wl.txt
done
timeout 300 python -m pytest tests/test_gpu_collectives.py tests/test_gpu_codec.py -x -q -m gpu 2>&1 | tail -3
