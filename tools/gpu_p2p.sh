#!/bin/bash
# peer-memory transport of the compressed collectives on N GPUs of one box (gpurun --gpus N)
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/allgather_p2p.py 2>&1 | grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" | tee gpurun_out/p2p_${N}gpu.txt
timeout 600 python -m pytest tests/test_gpu_collectives.py tests/test_gpu_codec.py -x -q -m gpu 2>&1 | tail -4
