#!/bin/bash
# 2-GPU check of both bench arms (torchrun, NCCL) and of the compressed collectives over NCCL
cd "$GRAFT_REPO_ROOT"
nvidia-smi topo -m 2>/dev/null | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench_ours_2gpu.json 2> gpurun_out/r2m_bench_ours_2gpu.err; cut -c1-200 gpurun_out/r2m_bench_ours_2gpu.json; tail -2 gpurun_out/r2m_bench_ours_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench_ref_2gpu.json 2> gpurun_out/r2m_bench_ref_2gpu.err; cut -c1-200 gpurun_out/r2m_bench_ref_2gpu.json; tail -2 gpurun_out/r2m_bench_ref_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/allgather_2gpu.py 2>&1 | grep -v Warning | tee gpurun_out/r2m_allgather_2gpu.txt
timeout 300 python -m pytest tests/test_gpu_collectives.py -x -q -m gpu 2>&1 | tail -2
