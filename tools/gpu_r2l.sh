#!/bin/bash
# round 2, step L: free-running-warps decoder (decode_fused=2) vs lease decoder (1); unpredicated refill load
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/walltime.py c3 "" "decode_fused=2" "decode_fused=0" 2>&1 | tee gpurun_out/r2l_wall_c3.txt
timeout 600 python -c "
from dietgpu_b200 import capi
capi.set_option('decode_fused', 2)
import pytest, sys
sys.exit(pytest.main(['tests/test_gpu_codec.py', 'tests/test_reference_parity.py', 'tests/test_gpu_host.py', '-x', '-q', '-m', 'gpu', '-p', 'no:cacheprovider']))
" > gpurun_out/r2l_pytest_free.log 2>&1; tail -3 gpurun_out/r2l_pytest_free.log
for wl in c4 c2 c3x1; do
timeout 300 python tools/walltime.py $wl "" "decode_fused=2" 2>&1 | tee gpurun_out/r2l_wall_$wl.txt
done
