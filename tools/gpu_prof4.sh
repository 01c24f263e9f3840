#!/bin/bash
cd "$GRAFT_REPO_ROOT"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:decodeKernel -s 2 -c 1 -o gpurun_out/prof4_decode_c3 -f python tools/prof_one.py c3 3 decode_warps=8 > gpurun_out/prof4_decode_c3.log 2>&1
$NCU -k regex:decodeKernel -s 2 -c 1 -o gpurun_out/prof4_decode_c2 -f python tools/prof_one.py c2 3 decode_warps=8 > gpurun_out/prof4_decode_c2.log 2>&1
