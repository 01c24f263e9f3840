#!/bin/bash
# round 2, step I: TMA-staged K1 beside a coder capped at 3 CTAs/SM
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/walltime.py c3 "" "stats_stage=1" "encode_k2_ctas=3" "stats_stage=1,encode_k2_ctas=3" "stats_stage=1,encode_k2_ctas=3,parts=8" "stats_stage=1,stats_stage_kb=16,encode_k2_ctas=3" "stats_stage=1,stats_stage_kb=64,encode_k2_ctas=3" "stats_stage=1,encode_k2_ctas=2,parts=8" 2>&1 | tee gpurun_out/r2i_wall_c3.txt
timeout 300 python tools/walltime.py c4 "" "stats_stage=1,encode_k2_ctas=3" "stats_stage=1,encode_k2_ctas=3,parts=8" 2>&1 | tee gpurun_out/r2i_wall_c4.txt
timeout 200 python tools/sweep.py c3 "stats_stage=1" 2>&1 | tee gpurun_out/r2i_sweep_c3.txt
