"""2-rank NCCL check of dietgpu_b200.all_gather_compressed (launch with torchrun --nproc-per-node 2):
bit-exact result, and device-timed compressed vs plain all-gather of the same bf16 shard."""
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
for dt, n in ((torch.bfloat16, 32 << 20), (torch.float32, 1 << 20), (torch.uint8, 3000001)):
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(n, generator=g, device=dev).to(dt) if dt.is_floating_point else \
        torch.randint(0, 50, (n,), generator=g, device=dev, dtype=torch.int32).to(dt)
    want = torch.empty(world * n, dtype=dt, device=dev)
    dist.all_gather_into_tensor(want, x)
    got = dg.all_gather_compressed(x)
    ok = torch.equal(got.view(torch.uint8), want.view(torch.uint8))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize(); dist.barrier()
    ev[0].record()
    for _ in range(5):
        dist.all_gather_into_tensor(want, x)
    ev[1].record()
    for _ in range(5):
        dg.all_gather_compressed(x)
    ev[2].record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{dt} n={n} world={world} bit_exact={ok} plain={ev[0].elapsed_time(ev[1]) / 5:.3f} ms "
              f"compressed={ev[1].elapsed_time(ev[2]) / 5:.3f} ms", flush=True)
dist.destroy_process_group()
