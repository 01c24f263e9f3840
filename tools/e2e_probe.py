"""End-to-end (pinned host -> archives on host -> pinned host) throughput of HostCodec for several group counts:
blocking calls and the software-pipelined async loop.  python tools/e2e_probe.py [workload] [groups ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import dietgpu_b200 as dg  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
groups = [int(g) for g in sys.argv[2:]] or [4, 8, 16, 32]
kind, batch, per, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
bench.bind_to_gpu_numa_node(0)
ts = bench.make_batch(torch, kind, batch, per, 1234, dev)
ub = sum(t.numel() * t.element_size() for t in ts)


def pinned_like(tensors):
    flat = torch.empty(sum(t.numel() * t.element_size() for t in tensors), dtype=torch.uint8, pin_memory=True)
    out, off = [], 0
    for t in tensors:
        nb = t.numel() * t.element_size()
        out.append(flat[off:off + nb].view(t.dtype).view(t.shape))
        off += nb
    return out


pin_in, pin_out = pinned_like(ts), pinned_like(ts)
for p, t in zip(pin_in, ts):
    p.copy_(t)
it = torch.int16 if kind != "bytes" else torch.uint8
for g in groups:
    hc = dg.HostCodec(kind != "bytes", pin_in, device=dev, groups=g)
    comp = [torch.empty((len(ts), hc.max_archive_bytes()), dtype=torch.uint8, pin_memory=True) for _ in range(2)]

    def sync_step():
        hs = hc.compress(pin_in, comp[0])
        hc.decompress([comp[0][i, :n] for i, n in enumerate(hs)], pin_out)

    def pipelined(n):
        pend = hc.compress_async(pin_in, comp[0])
        for i in range(n):
            hs = pend.finish()
            pd = hc.decompress_async([comp[i & 1][j, :k] for j, k in enumerate(hs)], pin_out)
            if i + 1 < n:
                pend = hc.compress_async(pin_in, comp[(i + 1) & 1])
                pend.finish()
            pd.finish()

    sync_step(); pipelined(2); torch.cuda.synchronize()
    n = 8
    t0 = time.perf_counter()
    for _ in range(n):
        sync_step()
    torch.cuda.synchronize()
    ts_ = time.perf_counter() - t0
    t0 = time.perf_counter()
    pipelined(2 * n)
    torch.cuda.synchronize()
    tp = time.perf_counter() - t0
    ok = all(torch.equal(a.view(it), b.view(it).to(dev)) for a, b in zip(ts, pin_out))
    print(f"{wl} groups={g}: sync {2 * ub * n / ts_ / 1e9:.1f} GB/s ({ts_ / n * 1e3:.2f} ms/step) | pipelined "
          f"{2 * ub * 2 * n / tp / 1e9:.1f} GB/s ({tp / (2 * n) * 1e3:.2f} ms/step) ok={ok}", flush=True)
    del hc, comp
