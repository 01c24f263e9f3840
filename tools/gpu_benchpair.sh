#!/bin/bash
# both default bench lines from one box
cd "$GRAFT_REPO_ROOT"
T=${1:-pair}
timeout 120 python bench.py --impl reference 2> gpurun_out/${T}_bench_ref.err | grep "^{" > gpurun_out/${T}_bench_ref.json; cut -c1-160 gpurun_out/${T}_bench_ref.json
timeout 120 python bench.py 2> gpurun_out/${T}_bench_ours.err | grep "^{" > gpurun_out/${T}_bench_ours.json; cut -c1-160 gpurun_out/${T}_bench_ours.json
