#!/bin/bash
# ncu captures of the hot kernels (one warm launch each), summaries land in gpurun_out/
set -x
cd "$GRAFT_REPO_ROOT"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:decodeKernel -s 2 -c 1 -o gpurun_out/prof_decode_c3 -f python tools/prof_one.py c3 3 > gpurun_out/prof_decode_c3.log 2>&1
$NCU -k regex:decodeKernel -s 2 -c 1 -o gpurun_out/prof_decode_c2 -f python tools/prof_one.py c2 3 > gpurun_out/prof_decode_c2.log 2>&1
$NCU -k regex:encodeKernel -s 2 -c 1 -o gpurun_out/prof_encode_c3 -f python tools/prof_one.py c3 3 > gpurun_out/prof_encode_c3.log 2>&1
$NCU -k regex:statsFloatKernel -s 2 -c 1 -o gpurun_out/prof_stats_c3 -f python tools/prof_one.py c3 3 > gpurun_out/prof_stats_c3.log 2>&1
$NCU -k regex:statsFloatKernel -s 2 -c 1 -o gpurun_out/prof_stats_c4 -f python tools/prof_one.py c4 3 > gpurun_out/prof_stats_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
