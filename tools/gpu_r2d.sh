#!/bin/bash
# round 2, step D: lease decoder + async HostCodec + new tests; full default bench line
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
for wl in c3 c2 c4 c3x1; do
timeout 300 python tools/walltime.py $wl "encode_fused=0" "encode_fused=0,decode_fused=0" 2>&1 | tee gpurun_out/r2d_wall_$wl.txt
done
( time timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err ) 2>&1 | tail -3; cut -c1-1500 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
