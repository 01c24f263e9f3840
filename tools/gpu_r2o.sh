#!/bin/bash
# r2o: ring prefetch issued after the warp barrier (formal WAR fix): tests, timing, sanitizer passes
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for wl in c3 c2 c4 c3x1; do
  WALL_GRAPH=0 timeout 300 python tools/walltime.py $wl "" "" 2>&1 | grep -v Warning
done | tee gpurun_out/r2o_wall.txt
{
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool ==="
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python tools/sanitize_small.py 2>&1 | grep -v "Warning\|warn" | tail -8
  echo "exit code: ${PIPESTATUS[0]}"
done
} > gpurun_out/sanitizer.txt 2>&1
cut -c1-200 gpurun_out/sanitizer.txt
