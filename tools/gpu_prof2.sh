#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:decodeKernel -s 2 -c 1 -o gpurun_out/prof2_decode_c3 -f python tools/prof_one.py c3 3 > gpurun_out/prof2_decode_c3.log 2>&1
$NCU -k regex:decodeKernel -s 2 -c 1 -o gpurun_out/prof2_decode_c3_l64 -f python tools/prof_one.py c3 3 decode_lut64=1 > gpurun_out/prof2_decode_c3_l64.log 2>&1
ls -la gpurun_out/*.ncu-rep
