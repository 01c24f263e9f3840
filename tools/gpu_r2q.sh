#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for wl in c3 c4; do
  WALL_GRAPH=0 timeout 300 python tools/walltime.py $wl "" "last_part_pct=50" "first_part_pct=130,last_part_pct=50" "last_part_pct=70" "first_part_pct=130" "parts=8,last_part_pct=50" "" 2>&1 | grep -v Warning
done | tee gpurun_out/r2q_wall2.txt
