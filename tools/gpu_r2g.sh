#!/bin/bash
# round 2, step G: defaults ring3 + parts 4; e2e group-count probe
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
timeout 300 python tools/walltime.py c3 "" "parts=8" "parts=2" "hist_slab_kb=113" "hist_slab_kb=113,parts=8" 2>&1 | tee gpurun_out/r2g_wall_c3.txt
timeout 300 python tools/walltime.py c4 "" "parts=8" 2>&1 | tee gpurun_out/r2g_wall_c4.txt
timeout 300 python tools/e2e_probe.py c3 4 8 16 32 2>&1 | tee gpurun_out/r2g_e2e_probe.txt
