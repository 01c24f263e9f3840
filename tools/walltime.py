"""Wall (event) time of encode / decode calls for option variants: python tools/walltime.py <wl> 'opt=v,...' ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dietgpu_b200 import capi  # noqa: E402

wl = sys.argv[1]
kind, batch, per, desc = bench.WORKLOADS[wl]
ts = bench.make_batch(torch, kind, batch, per, 1234, torch.device("cuda", 0))
ub = sum(t.numel() * t.element_size() for t in ts)
it = torch.int16 if kind != "bytes" else torch.uint8
defaults = {}
sizes0 = None
for v in sys.argv[2:] or [""]:
    for k, val in defaults.items():  # every variant starts from the defaults
        capi.set_option(k, val)
    for kv in [x for x in v.split(",") if x]:
        k, val = kv.split("=")
        defaults.setdefault(k, capi.get_option(k))
        capi.set_option(k, int(val))
    c = bench.OursCodec(torch, kind, ts)
    c.encode(); c.bind_rows(); c.decode(); torch.cuda.synchronize()
    ok = all(torch.equal(a.view(it), b.view(it)) for a, b in zip(ts, c.outs))
    sz = c.sizes.cpu().tolist()
    sizes0 = sizes0 or sz
    ok = ok and sz == sizes0  # archive sizes do not depend on the kernel variant
    for _ in range(3):
        c.encode(); c.decode()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    n = 30
    e[0].record()
    for _ in range(n):
        c.encode()
    e[1].record()
    for _ in range(n):
        c.decode()
    e[2].record()
    torch.cuda.synchronize()
    te, td = e[0].elapsed_time(e[1]) / n * 1e3, e[1].elapsed_time(e[2]) / n * 1e3
    # the bench's step: encode and decode alternate on one stream
    e[0].record()
    for _ in range(n):
        c.encode(); c.decode()
    e[1].record()
    torch.cuda.synchronize()
    ts_ = e[0].elapsed_time(e[1]) / n * 1e3
    # the same step captured once into a CUDA graph and replayed (possible when the member table travels in the
    # kernel parameters: nothing of the call reads host memory later)
    tg = float("nan")
    if os.environ.get("WALL_GRAPH", "1") == "1":
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                c.encode(); c.decode()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e[0].record()
            for _ in range(n):
                g.replay()
            e[1].record()
            torch.cuda.synchronize()
            tg = e[0].elapsed_time(e[1]) / n * 1e3
            ok = ok and all(torch.equal(a.view(it), b.view(it)) for a, b in zip(ts, c.outs))
        except Exception as ex:  # noqa: BLE001
            print("graph capture failed:", str(ex)[:200])
    print(f"{wl} [{v}] ok={ok} enc {te:.1f}us {ub / te / 1e3:.0f} GB/s | dec {td:.1f}us {ub / td / 1e3:.0f} GB/s | both {2 * ub / (te + td) / 1e3:.0f} GB/s"
          f" | step {ts_:.1f}us {2 * ub / ts_ / 1e3:.0f} GB/s | graph step {tg:.1f}us {2 * ub / tg / 1e3:.0f} GB/s", flush=True)
    del c
