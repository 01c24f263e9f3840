"""Experiment: does the HBM-bound stats kernel (K1) of one sub-batch overlap with the issue-bound encode
kernel (K2) of another when sub-batches are submitted on different streams?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dietgpu_b200 import capi  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
kind, batch, per, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
ts = bench.make_batch(torch, kind, batch, per, 1234, dev)
ub = sum(t.numel() * t.element_size() for t in ts)
for opts in sys.argv[2:] or [""]:
    for kv in [x for x in opts.split(",") if x]:
        k, v = kv.split("=")
        if k != "parts":
            capi.set_option(k, int(v))
    for parts in (1, 2, 4, 8):
        chunk = batch // parts
        codecs = [bench.OursCodec(torch, kind, ts[i * chunk:(i + 1) * chunk]) for i in range(parts)]
        streams = [torch.cuda.Stream() for _ in range(parts)]
        def enc_all():
            for c, s in zip(codecs, streams):
                with torch.cuda.stream(s):
                    c.encode()
        def dec_all():
            for c, s in zip(codecs, streams):
                with torch.cuda.stream(s):
                    c.decode()
        enc_all(); torch.cuda.synchronize()
        for c in codecs:
            c.bind_rows()
        dec_all(); torch.cuda.synchronize()
        ok = all(torch.equal(a.view(torch.int16), b.view(torch.int16)) for c in codecs for a, b in zip(c.ts, c.outs))
        res = {}
        for name, fn in (("enc", enc_all), ("dec", dec_all)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 10
            for _ in range(n):
                fn()
                for s in streams:
                    torch.cuda.current_stream().wait_stream(s)
                for s in streams:
                    s.wait_stream(torch.cuda.current_stream())
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / n * 1e3
        print(f"{wl} [{opts}] parts={parts} ok={ok} enc={res['enc']:.1f}us ({ub / res['enc'] / 1e3:.0f} GB/s) "
              f"dec={res['dec']:.1f}us ({ub / res['dec'] / 1e3:.0f} GB/s)", flush=True)
