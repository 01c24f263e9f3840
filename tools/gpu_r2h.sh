#!/bin/bash
# round 2, step H: TMA bulk input ring in the coder (default) vs cp.async; device API example
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
for wl in c3 c4 c2; do
timeout 300 python tools/walltime.py $wl "" "encode_fused=1" 2>&1 | sed 's/^/tma: /' | tee gpurun_out/r2h_wall_${wl}_tma.txt
DIETGPU_B200_LIB=$PWD/dietgpu_b200/libdietgpu_b200_notma.so timeout 300 python tools/walltime.py $wl "" "encode_fused=1" 2>&1 | sed 's/^/cp.async: /' | tee gpurun_out/r2h_wall_${wl}_cpasync.txt
done
timeout 200 python tools/sweep.py c3 "" 2>&1 | tee gpurun_out/r2h_sweep_c3.txt
