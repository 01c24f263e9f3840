#!/bin/bash
# round-end evidence run (1 GPU): tests, both bench arms, ncu launch list, ncu full captures of the hot kernels
# usage: bash tools/gpu_final.sh [tag]   (outputs gpurun_out/<tag>_*)
cd "$GRAFT_REPO_ROOT"
T=${1:-final}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err; cut -c1-300 gpurun_out/${T}_bench_ours.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-detail > gpurun_out/${T}_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:decodeFusedKernel -s 2 -c 1 -o gpurun_out/${T}_decode_c3 -f python tools/prof_one.py c3 3 > /dev/null 2>&1
timeout 300 $NCU -k regex:encodeKernelFast -s 2 -c 1 -o gpurun_out/${T}_encode_c3 -f python tools/prof_one.py c3 3 parts=1 > /dev/null 2>&1
timeout 300 $NCU -k regex:statsFloatKernel -s 2 -c 1 -o gpurun_out/${T}_stats_c3 -f python tools/prof_one.py c3 3 parts=1 > /dev/null 2>&1
timeout 300 $NCU -k regex:decodeFusedKernel -s 2 -c 1 -o gpurun_out/${T}_decode_c2 -f python tools/prof_one.py c2 3 > /dev/null 2>&1
timeout 300 $NCU -k regex:encodeKernelFast -s 2 -c 1 -o gpurun_out/${T}_encode_c2 -f python tools/prof_one.py c2 3 parts=1 > /dev/null 2>&1
{ for tool in memcheck racecheck; do echo "=== compute-sanitizer --tool $tool: device-level API example kernels (tests/test_gpu_device_api.py) ==="; timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 6 python -m pytest tests/test_gpu_device_api.py -x -q -m gpu 2>&1 | grep -v "Warning: \|warn\|Host Frame\|host backtrace" | tail -6; echo "exit code: ${PIPESTATUS[0]}"; done; } > gpurun_out/${T}_sanitizer_device_api.txt 2>&1
ls -la gpurun_out/${T}_*
