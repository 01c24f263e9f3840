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
ws = dg.PeerWorkspace((world + 1) * (int(max(sizes_mib) * (1 << 20) * 1.25) + (1 << 20)))
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
    t_peer = timed(lambda: dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, peer_mode="pull"))
    gotp = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, peer_mode="push")
    ok = ok and torch.equal(gotp.view(torch.int16), want.view(torch.int16))
    t_push = {}
    for st_ in (1, 2, 4):
        t_push[st_] = timed(lambda: dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, stages=st_, peer_mode="push"))
    got3 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, peer_mode="direct")
    ok = ok and torch.equal(got3.view(torch.int16), want.view(torch.int16))
    t_direct = timed(lambda: dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, peer_mode="direct"))
    # the same call without the Python / ctypes launch path: two calls (both halves of the workspace, so the
    # one-barrier-per-call ordering argument holds across replays) captured into one CUDA graph
    t_graph = t_graph_direct = float("nan")
    t_graph_push = {}
    try:
        torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            o1 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, peer_mode="pull")
            o2 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, peer_mode="pull")
        t_graph = timed(gr.replay) / 2
        ok = ok and torch.equal(o1.view(torch.int16), want.view(torch.int16)) and torch.equal(o2.view(torch.int16), want.view(torch.int16))
        torch.cuda.synchronize(); dist.barrier()
        gd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gd):
            dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, peer_mode="direct")
            dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, peer_mode="direct")
        t_graph_direct = timed(gd.replay) / 2
        t_graph_push = {}
        for st_ in (1, 2, 4):
            torch.cuda.synchronize(); dist.barrier()
            gp = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gp):
                p1 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, stages=st_, peer_mode="push")
                p2 = dg.all_gather_compressed(x, members=members, peer=ws, temp_mem=temp, check=False, stages=st_, peer_mode="push")
            t_graph_push[st_] = timed(gp.replay) / 2
            ok = ok and torch.equal(p1.view(torch.int16), want.view(torch.int16)) and torch.equal(p2.view(torch.int16), want.view(torch.int16))
    except Exception as ex:  # noqa: BLE001
        if rank == 0:
            print("graph capture failed:", str(ex)[:300], flush=True)
    # the pieces, each alone and graph-replayed: decode of one peer's archives read over NVLink vs the same
    # archives read from local memory
    def graphed(fn):
        torch.cuda.synchronize(); dist.barrier()
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            fn()
        return timed(gg.replay)

    per = n // members
    cols = dg.collectives._archive_cols(True, torch.bfloat16, per)
    peer_rank = (rank + 1) % world
    off = 0  # both halves hold this size's archives after the calls above
    rows_peer = ws.view(peer_rank, off, members * cols).view(members, cols)
    rows_local = rows_peer.clone()
    outs = [torch.empty(per, dtype=torch.bfloat16, device=dev) for _ in range(members)]
    t_dec_peer = graphed(lambda: dg.decompress_data(True, [rows_peer[i] for i in range(members)], outs, False, temp))
    t_dec_local = graphed(lambda: dg.decompress_data(True, [rows_local[i] for i in range(members)], outs, False, temp))
    t_copy_peer = graphed(lambda: rows_local.copy_(rows_peer))
    t_pull = {}
    for ctas in (16, 32, 64, 128):
        dg.capi.set_option("pull_ctas", ctas)
        t_pull[ctas] = graphed(lambda: dg.ops.pull_archives(True, [rows_peer[i] for i in range(members)],
                                                            [rows_local[i] for i in range(members)], torch.bfloat16))
    dg.capi.set_option("pull_ctas", 64)
    inbox_peer = ws.view(peer_rank, (1 + rank) * members * cols, members * cols).view(members, cols)  # where push mode writes
    t_pushmv = graphed(lambda: dg.ops.pull_archives(True, [rows_local[i] for i in range(members)],
                                                    [inbox_peer[i] for i in range(members)], torch.bfloat16))
    # the pieces of the peer transport, each alone
    comp, csz, _ = dg.compress_data(True, [x[i * (n // members):(i + 1) * (n // members)] for i in range(members)], False, temp)
    t_enc = timed(lambda: dg.compress_data(True, [x[i * (n // members):(i + 1) * (n // members)] for i in range(members)], False, temp, comp, csz))
    ratio = csz.sum().item() / (2 * n)
    if rank == 0:
        bus = (world - 1) * 2 * n / 1e6  # MB every rank receives
        print(f"all_gather bf16 {mib} MiB/rank world={world} bit_exact={ok} ratio={ratio:.3f} | plain NCCL {t_plain:.3f} ms "
              f"({bus / t_plain:.0f} GB/s in) | compressed over NCCL {t_nccl:.3f} ms | compressed, peer pull "
              f"(mover + local decode, pipelined) {t_peer:.3f} ms ({bus / t_peer:.0f} GB/s in, {t_plain / t_peer:.2f}x plain), CUDA-graph replay "
              f"{t_graph:.3f} ms ({t_plain / t_graph:.2f}x plain) | compressed, decoder reads peer memory directly {t_direct:.3f} ms, graph replay "
              f"{t_graph_direct:.3f} ms ({t_plain / t_graph_direct:.2f}x plain) | compressed, PUSH (mover writes into the peers' inboxes + flags; stages: host-launched / graph replay ms): "
              + ", ".join(f"{k}: {t_push[k]:.3f} / {t_graph_push.get(k, float('nan')):.3f} ({t_plain / t_graph_push.get(k, float('nan')):.2f}x plain)" for k in t_push)
              + f" | encode alone (host-launched) {t_enc:.3f} ms | decode of one peer's {members} archives: over NVLink {t_dec_peer:.3f} ms "
              f"({ratio * 2 * n / 1e6 / t_dec_peer:.0f} GB/s pulled), from local memory {t_dec_local:.3f} ms; plain copy of the padded rows over NVLink "
              f"{t_copy_peer:.3f} ms ({members * cols / 1e6 / t_copy_peer:.0f} GB/s); archive mover (exact bytes) by grid size: "
              + ", ".join(f"{c} CTAs {t:.3f} ms ({ratio * 2 * n / 1e6 / t:.0f} GB/s)" for c, t in t_pull.items())
              + f"; same mover WRITING local archives into peer memory (64 CTAs) {t_pushmv:.3f} ms ({ratio * 2 * n / 1e6 / t_pushmv:.0f} GB/s)",
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
