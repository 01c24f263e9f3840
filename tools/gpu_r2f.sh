#!/bin/bash
# round 2, step F: K1/K2 tuning after the coder took over the split; lease decoder with claim-first; e2e with coalesced copies
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python tools/walltime.py c3 "" "parts=1" "parts=4" "hist_slab_kb=32" "hist_slab_kb=128" "hist_slab_kb=113" "encode_slot_words=1280" "decode_fused=0" 2>&1 | tee gpurun_out/r2f_wall_c3.txt
DIETGPU_B200_LIB=$PWD/dietgpu_b200/libdietgpu_b200_ring3.so timeout 300 python tools/walltime.py c3 "" "encode_slot_words=1280" "encode_slot_words=1024" 2>&1 | sed 's/^/ring3: /' | tee gpurun_out/r2f_wall_c3_ring3.txt
DIETGPU_B200_LIB=$PWD/dietgpu_b200/libdietgpu_b200_ring3.so timeout 300 python tools/walltime.py c4 "" "encode_slot_words=1280" 2>&1 | sed 's/^/ring3: /' | tee gpurun_out/r2f_wall_c4_ring3.txt
for wl in c4 c2 c3x1; do
timeout 300 python tools/walltime.py $wl "" "decode_fused=0" 2>&1 | tee gpurun_out/r2f_wall_$wl.txt
done
timeout 600 python bench.py --no-detail --steps 20 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print({k:(d[k]['value'] if isinstance(d.get(k),dict) else d.get(k)) for k in ('value','e2e','e2e_sync','e2e_plain')})
PY
tail -3 gpurun_out/r2f_bench.err
