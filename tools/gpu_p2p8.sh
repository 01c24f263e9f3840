#!/bin/bash
# peer-memory transports of the compressed collectives on all N GPUs of the box
cd "$GRAFT_REPO_ROOT"
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -12 | cut -c1-120
SIZES_MIB=${SIZES_MIB:-64,256} REPS=5 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 tools/allgather_p2p.py 2>&1 | grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" | tee gpurun_out/p2p_${N}gpu.txt
