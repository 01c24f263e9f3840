#!/bin/bash
# round 2, step B: why does the fused encoder not beat K1+K2?  ncu on it, variants, PCIe duplex
cd "$GRAFT_REPO_ROOT"
timeout 120 python tools/pcie_duplex.py 2>&1 | tee gpurun_out/r2b_pcie_duplex.txt
timeout 300 python tools/walltime.py c3 "" "encode_wide_table=0" "encode_fused=0,encode_wide_table=0" "decode_chunk_blocks=8" "decode_chunk_blocks=32" "decode_fused=0" "encode_slot_words=1280" 2>&1 | tee gpurun_out/r2b_wall_c3.txt
DIETGPU_B200_LIB=$PWD/dietgpu_b200/libdietgpu_b200_u8.so timeout 300 python tools/walltime.py c3 "" "encode_fused=0" "encode_wide_table=0" 2>&1 | sed 's/^/u8: /' | tee gpurun_out/r2b_wall_c3_u8.txt
timeout 200 python tools/walltime.py c2 "" "decode_fused=0" "decode_chunk_blocks=8" 2>&1 | tee gpurun_out/r2b_wall_c2.txt
timeout 200 python tools/walltime.py c4 "" "decode_fused=0" "decode_chunk_blocks=8" "encode_wide_table=1" 2>&1 | tee gpurun_out/r2b_wall_c4.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:encodeFusedKernel -s 2 -c 1 -o gpurun_out/r2b_fused_c3 -f python tools/prof_one.py c3 3 > gpurun_out/r2b_ncu1.log 2>&1; tail -2 gpurun_out/r2b_ncu1.log
timeout 300 $NCU -k regex:decodeFusedKernel -s 2 -c 1 -o gpurun_out/r2b_decfused_c3 -f python tools/prof_one.py c3 3 > gpurun_out/r2b_ncu2.log 2>&1; tail -2 gpurun_out/r2b_ncu2.log
ls -la gpurun_out/r2b_*
