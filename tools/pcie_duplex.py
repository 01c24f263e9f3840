"""PCIe link micro-benchmark: pinned host memory, 64 MiB chunks, H2D alone, D2H alone, both at once on
two streams (is the link full duplex on this box?), and both at once with the process bound to the
CPUs of each NUMA node in turn (first-touch places the pinned pages on that node).
    python tools/pcie_duplex.py [gpu_index]
"""
import glob
import os
import sys

import torch

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.cuda.set_device(dev)
CH = 64 << 20
N = 8


def node_cpus():
    nodes = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*/cpulist")):
        n = int(p.split("node")[-1].split("/")[0])
        cpus = []
        for part in open(p).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus += list(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        nodes[n] = cpus
    return nodes


def bench(label):
    hin = [torch.empty(CH, dtype=torch.uint8).pin_memory() for _ in range(N)]
    hout = [torch.empty(CH, dtype=torch.uint8).pin_memory() for _ in range(N)]
    for h in hin + hout:
        h.fill_(1)  # first touch
    din = [torch.empty(CH, dtype=torch.uint8, device="cuda") for _ in range(N)]
    dout = [torch.ones(CH, dtype=torch.uint8, device="cuda") for _ in range(N)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, reps=4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            for i in range(N):
                if up:
                    with torch.cuda.stream(s1):
                        din[i].copy_(hin[i], non_blocking=True)
                if down:
                    with torch.cuda.stream(s2):
                        hout[i].copy_(dout[i], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return reps * N * CH / (e0.elapsed_time(e1) / 1e3) / 1e9

    run(True, True, 1)
    u, d, b = run(True, False), run(False, True), run(True, True)
    print(f"{label}: H2D alone {u:.1f} GB/s | D2H alone {d:.1f} GB/s | both at once {b:.1f} + {b:.1f} = {2 * b:.1f} GB/s total", flush=True)


print(torch.cuda.get_device_name(dev), "gpu", dev)
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(dev)
    print("pcie gen", nv.nvmlDeviceGetCurrPcieLinkGeneration(h), "width", nv.nvmlDeviceGetCurrPcieLinkWidth(h),
          "| gpu numa node:", open(f"/sys/bus/pci/devices/{nv.nvmlDeviceGetPciInfo(h).busId.lower()[4:] if len(nv.nvmlDeviceGetPciInfo(h).busId) > 12 else nv.nvmlDeviceGetPciInfo(h).busId.lower()}/numa_node").read().strip())
except Exception as e:  # noqa: BLE001
    print("nvml/sysfs:", e)
bench("unbound")
all_cpus = os.sched_getaffinity(0)
for node, cpus in node_cpus().items():
    usable = set(cpus) & all_cpus
    if not usable:
        continue
    os.sched_setaffinity(0, usable)
    bench(f"bound to node {node} ({len(usable)} cpus)")
os.sched_setaffinity(0, all_cpus)
