"""Runs a few encode/decode iterations of one workload so `ncu -k regex:... -s N -c 1` can capture a
warm launch.  Usage: python tools/prof_one.py <workload> [iters] [option=value ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dietgpu_b200 import capi  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    capi.set_option(k, int(v))
kind, batch, per, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
ts = bench.make_batch(torch, kind, batch, per, 1234, dev)
codec = bench.OursCodec(torch, kind, ts)
for _ in range(iters):
    codec.encode()
    codec.bind_rows()
    codec.decode()
torch.cuda.synchronize()
print("done", desc)
