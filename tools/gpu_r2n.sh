#!/bin/bash
# r2n: member table in the kernel parameters (A/B against the upload), CUDA-graph replay of a step, GPU tests
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for wl in c3 c2 c4 c3x1; do
  timeout 300 python tools/walltime.py $wl "" "inline_members=0" "" "inline_members=0" 2>&1 | grep -v Warning
done | tee gpurun_out/r2n_wall.txt
