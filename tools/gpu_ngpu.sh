#!/bin/bash
# N-GPU check of both bench arms under torchrun, as the driver launches them
cd "$GRAFT_REPO_ROOT"
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/bench_ours_${N}gpu.err | grep "^{" > gpurun_out/bench_ours_${N}gpu.json; cut -c1-220 gpurun_out/bench_ours_${N}gpu.json; tail -2 gpurun_out/bench_ours_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --impl reference --gpus $N --steps 20 --warmup 3 2> gpurun_out/bench_ref_${N}gpu.err | grep "^{" > gpurun_out/bench_ref_${N}gpu.json; cut -c1-220 gpurun_out/bench_ref_${N}gpu.json; tail -2 gpurun_out/bench_ref_${N}gpu.err
