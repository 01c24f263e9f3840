#!/bin/bash
cd "$GRAFT_REPO_ROOT"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:encodeKernelFast -s 2 -c 1 -o gpurun_out/prof7_encode_c3 -f python tools/prof_one.py c3 3 parts=1 > gpurun_out/prof7_encode_c3.log 2>&1
