"""Compressed all-gather: the use the reference names for this codec ("compress before a collective",
README.md:72, 103-104) and ships no code for.

Each rank compresses its shard, the ranks exchange the variable-size archives, every rank decompresses
everything.  The exchange is two collectives: the archive sizes (a few bytes), then one
all_gather_into_tensor of each rank's packed archives padded to the largest rank total, so the wire
carries about `ratio` x the raw bytes (0.67 for bf16 activations) instead of all of them.
The codec work is the ordinary operator path (dietgpu_b200.ops -> C ABI -> sm_100a kernels) on the
caller's device; there is no CPU fallback for it.  `exchange_archives` itself is plain
torch.distributed plumbing and runs on any backend / device the process group supports.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops

_ALIGN = 16  # archives are 16 B aligned in the packed buffer (compressed buffers must be, ans/GpuANSEncode.cu:19-21)


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def pack_offsets(sizes: Sequence[int]) -> Tuple[List[int], int]:
    """Offsets of archives of the given byte sizes in one packed buffer, each 16 B aligned; total bytes."""
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += _round_up(int(s), _ALIGN)
    return offs, o


def exchange_archives(rows: Sequence[torch.Tensor], group=None) -> List[List[torch.Tensor]]:
    """All ranks contribute the same NUMBER of archives (uint8 1-D tensors of any sizes, one device).
    Returns, for every rank in rank order, the list of that rank's archives (views into one gathered
    buffer on the same device).  Two collectives: sizes, then the padded payload."""
    n = len(rows)
    if n == 0:
        raise ValueError("exchange_archives: empty contribution")
    dev = rows[0].device
    for r in rows:
        if r.dtype != torch.uint8 or r.dim() != 1 or r.device != dev:
            raise ValueError("exchange_archives: archives must be uint8 1-D tensors on one device")
    world = _world(group)
    local_sizes = torch.tensor([r.numel() for r in rows], dtype=torch.int32, device=dev)
    if world == 1:
        return [[r for r in rows]]
    all_sizes = torch.empty(world * n, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_sizes, local_sizes, group=group)
    sizes = all_sizes.cpu().view(world, n).tolist()           # the one host sync of the exchange
    totals = [pack_offsets(s)[1] for s in sizes]
    width = max(max(totals), _ALIGN)
    rank = dist.get_rank(group)
    send = torch.zeros(width, dtype=torch.uint8, device=dev)
    offs, _ = pack_offsets(sizes[rank])
    for r, o in zip(rows, offs):
        send[o:o + r.numel()].copy_(r)
    recv = torch.empty(world * width, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    out: List[List[torch.Tensor]] = []
    for w in range(world):
        offs, _ = pack_offsets(sizes[w])
        base = w * width
        out.append([recv[base + o: base + o + s] for o, s in zip(offs, sizes[w])])
    return out


def all_gather_compressed(t: torch.Tensor, group=None, members: int = 8, checksum: bool = False,
                          temp_mem: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Every rank contributes the CUDA tensor `t` (same shape and dtype on every rank; fp16 / bf16 / fp32 go
    through the float codec, anything else through the byte codec) and receives the concatenation
    [world * t.numel()] in rank order, bit-exact.  `members` = archives per rank (the codec's parallelism
    comes from blocks, so a handful is enough; more only helps the sub-batch overlap on very large shards)."""
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("all_gather_compressed: contiguous CUDA tensor expected (no CPU fallback)")
    as_float = t.dtype in (torch.float16, torch.bfloat16, torch.float32)
    flat = t.reshape(-1) if as_float else t.reshape(-1).view(torch.uint8)
    n = flat.numel()
    members = max(1, min(int(members), max(1, n // 4096)))
    # equal member lengths (multiples of 8 elements keep every member 16 B aligned); the last takes the rest
    per = (n // members) // 8 * 8 if members > 1 else n
    bounds = [i * per for i in range(members)] + [n]
    chunks = [flat[bounds[i]:bounds[i + 1]] for i in range(members)]
    comp, sizes, _ = ops.compress_data(as_float, chunks, checksum, temp_mem)
    hs = sizes.cpu().tolist()
    rows = [comp[i, :hs[i]] for i in range(members)]
    gathered = exchange_archives(rows, group)
    world = len(gathered)
    out = torch.empty(world * n, dtype=flat.dtype, device=t.device)
    ins, outs = [], []
    for w in range(world):
        for i in range(members):
            ins.append(gathered[w][i])
            outs.append(out[w * n + bounds[i]: w * n + bounds[i + 1]])
    # gathered views are 16 B aligned inside the receive buffer; the decoder needs nothing else
    status = torch.zeros(len(ins), dtype=torch.uint8, device=t.device)
    ops.decompress_data(as_float, ins, outs, checksum, temp_mem, status)
    if not bool(status.all()):
        raise RuntimeError("all_gather_compressed: a received archive failed to decode")
    return out if as_float else out.view(t.dtype)
