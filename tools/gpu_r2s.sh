#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for wl in c2 c2p11 c3; do
  WALL_GRAPH=0 timeout 300 python tools/walltime.py $wl "" "decode_warps=8" "" 2>&1 | grep -v Warning
done | tee gpurun_out/r2s_wall.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 6 python tools/sanitize_small.py 2>&1 | grep -v "Warning: \|warn\|Host Frame" | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 6 python tools/sanitize_small.py 2>&1 | grep -v "Warning: \|warn\|Host Frame" | tail -4
