"""N-rank check and timing of the peer-memory transport of the compressed collectives (launch with torchrun
--nproc-per-node N): every rank encodes into a peer-mapped buffer, one device-side barrier, and the decode
kernel pulls the peers' archives over NVLink.  Compared with a plain NCCL all-gather / all-to-all of the same
data and with the NCCL transport of the compressed collective.  Device-timed, max over ranks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dietgpu_b200 as dg  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
REPS = int(os.environ.get("REPS", "10"))


def timed(fn, reps=REPS):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


sizes_mib = [int(x) for x in os.environ.get("SIZES_MIB", "16,64,256").split(",")]
ws = dg.PeerWorkspace(int(max(sizes_mib) * (1 << 20) * 1.1) + (1 << 20))
temp = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mib in sizes_mib:
    n = mib * (1 << 20) // 2
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(n, generator=g, device=dev).to(torch.bfloat16)
    want = torch.empty(world * n, dtype=torch.bfloat16, device=dev)
    dist.all_gather_into_tensor(want, x)
    members = 16 if mib >= 64 else 8
    got = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp)
    ok = torch.equal(got.view(torch.int16), want.view(torch.int16))
    got2 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp)  # the other half of the workspace
    ok = ok and torch.equal(got2.view(torch.int16), want.view(torch.int16))
    t_plain = timed(lambda: dist.all_gather_into_tensor(want, x))
    t_nccl = timed(lambda: dg.all_gather_compressed(x, members=members, temp_mem=temp), reps=3)
    t_peer = timed(lambda: dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False))
    # the same call without the Python / ctypes launch path: two calls (both halves of the workspace, so the
    # one-barrier-per-call ordering argument holds across replays) captured into one CUDA graph
    t_graph = float("nan")
    try:
        torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            o1 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False)
            o2 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False)
        t_graph = timed(gr.replay) / 2
        ok = ok and torch.equal(o1.view(torch.int16), want.view(torch.int16)) and torch.equal(o2.view(torch.int16), want.view(torch.int16))
    except Exception as ex:  # noqa: BLE001
        if rank == 0:
            print("graph capture failed:", str(ex)[:300], flush=True)
    # the pieces of the peer transport, each alone
    comp, csz, _ = dg.compress_data(True, [x[i * (n // members):(i + 1) * (n // members)] for i in range(members)], False, temp)
    t_enc = timed(lambda: dg.compress_data(True, [x[i * (n // members):(i + 1) * (n // members)] for i in range(members)], False, temp, comp, csz))
    ratio = csz.sum().item() / (2 * n)
    if rank == 0:
        bus = (world - 1) * 2 * n / 1e6  # MB every rank receives
        print(f"all_gather bf16 {mib} MiB/rank world={world} bit_exact={ok} ratio={ratio:.3f} | plain NCCL {t_plain:.3f} ms "
              f"({bus / t_plain / 1e3:.0f} GB/s in) | compressed over NCCL {t_nccl:.3f} ms | compressed, peer pull "
              f"{t_peer:.3f} ms ({bus / t_peer / 1e3:.0f} GB/s in, {t_plain / t_peer:.2f}x plain) | same, CUDA-graph replay {t_graph:.3f} ms "
              f"({t_plain / t_graph:.2f}x plain) | encode alone {t_enc:.3f} ms",
              flush=True)

# all-to-all: every rank sends a different chunk to every rank
for mib in sizes_mib:
    m = mib * (1 << 20) // 2 // world
    g = torch.Generator(device=dev).manual_seed(500 + rank)
    chunks = [torch.randn(m, generator=g, device=dev).to(torch.bfloat16) for _ in range(world)]
    send = torch.cat(chunks)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    got = dg.all_to_all_compressed(chunks, peer=ws, temp_mem=temp)
    ok = all(torch.equal(got[s].view(torch.int16), recv[s * m:(s + 1) * m].view(torch.int16)) for s in range(world))
    t_plain = timed(lambda: dist.all_to_all_single(recv, send))
    t_peer = timed(lambda: dg.all_to_all_compressed(chunks, peer=ws, temp_mem=temp, check=False))
    if rank == 0:
        print(f"all_to_all bf16 {mib} MiB/rank world={world} bit_exact={ok} | plain NCCL {t_plain:.3f} ms | compressed, peer pull "
              f"{t_peer:.3f} ms ({t_plain / t_peer:.2f}x plain)", flush=True)
dist.destroy_process_group()
