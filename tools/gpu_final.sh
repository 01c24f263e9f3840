#!/bin/bash
# round-end evidence run (1 GPU): tests, both bench arms, ncu launch list, ncu full captures
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
timeout 400 python bench.py --all-workloads > gpurun_out/final_bench_ours.json 2> gpurun_out/final_bench_ours.err; cat gpurun_out/final_bench_ours.json | cut -c1-400
timeout 400 python bench.py --impl reference --all-workloads --no-cpu > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cat gpurun_out/final_bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu > gpurun_out/final_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:decodeKernel -s 2 -c 1 -o gpurun_out/final_decode_c3 -f python tools/prof_one.py c3 3 parts=1 > /dev/null 2>&1
$NCU -k regex:encodeKernelFast -s 2 -c 1 -o gpurun_out/final_encode_c3 -f python tools/prof_one.py c3 3 parts=1 > /dev/null 2>&1
$NCU -k regex:statsFloatKernel -s 2 -c 1 -o gpurun_out/final_stats_c3 -f python tools/prof_one.py c3 3 parts=1 > /dev/null 2>&1
ls -la gpurun_out/final_*
python tools/walltime.py c3 "decode_slot_words=2048" "decode_slot_words=1536" "decode_slot_words=0" > gpurun_out/final_walltime.txt 2>&1; tail -3 gpurun_out/final_walltime.txt
