#!/bin/bash
cd "$GRAFT_REPO_ROOT"
SAN_QUICK=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 12 python tools/sanitize_small.py 2>&1 | grep -v "Warning: \|warn" | head -150 > gpurun_out/racecheck_detail.txt
tail -3 gpurun_out/racecheck_detail.txt
