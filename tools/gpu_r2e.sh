#!/bin/bash
# round 2, step E: coder reads the raw float words (statistics pass = pure read)
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
for wl in c3 c4 c3x1 c2; do
timeout 300 python tools/walltime.py $wl "" "encode_fused=1" "encode_fused=1,fused_stage=0" "encode_wide_table=0" "decode_fused=0" 2>&1 | tee gpurun_out/r2e_wall_$wl.txt
done
timeout 200 python tools/sweep.py c3 "" 2>&1 | tee gpurun_out/r2e_sweep_c3.txt
