#!/bin/bash
# compute-sanitizer passes over small round trips of every kernel flavour -> gpurun_out/sanitizer.txt
cd "$GRAFT_REPO_ROOT"
python tools/sanitize_small.py 2>&1 | tail -2
{
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== compute-sanitizer --tool $tool ==="
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python tools/sanitize_small.py 2>&1 | grep -v "Warning\|warn" | tail -12
  echo "exit code: ${PIPESTATUS[0]}"
done
} > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt | cut -c1-200
