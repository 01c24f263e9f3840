#!/bin/bash
# round 2, step C: parity suite on the staged fused encoder + K1 diet; timings on all workloads
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
for wl in c3 c2 c4 c3x1; do
timeout 300 python tools/walltime.py $wl "" "fused_stage=0" "encode_fused=0" "encode_slot_words=1280" "fused_chunk_blocks=8" 2>&1 | tee gpurun_out/r2c_wall_$wl.txt
done
