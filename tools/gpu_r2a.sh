#!/bin/bash
# round 2, step A: parity suite + first timings of the fused encoder
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
timeout 300 python tools/walltime.py c3 "" "encode_fused=0" "fused_stats_every=2" "fused_stats_every=8" "fused_stats_every=0" "fused_chunk_blocks=8" "fused_chunk_blocks=32" "hist_slab_kb=128" "hist_slab_kb=32" 2>&1 | tee gpurun_out/r2a_wall_c3.txt
timeout 200 python tools/walltime.py c2 "" "encode_fused=0" 2>&1 | tee gpurun_out/r2a_wall_c2.txt
timeout 200 python tools/walltime.py c4 "" "encode_fused=0" 2>&1 | tee gpurun_out/r2a_wall_c4.txt
timeout 200 python tools/walltime.py c3x1 "" "encode_fused=0" 2>&1 | tee gpurun_out/r2a_wall_c3x1.txt
