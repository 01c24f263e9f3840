#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
WALL_GRAPH=0 timeout 300 python tools/walltime.py c3 "" "" 2>&1 | grep -v Warning
{
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool ==="
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 6 python tools/sanitize_small.py 2>&1 | grep -v "Warning: \|warn\|Host Frame\|host backtrace" | tail -30
  echo "exit code: ${PIPESTATUS[0]}"
done
} > gpurun_out/sanitizer.txt 2>&1
cut -c1-250 gpurun_out/sanitizer.txt
