"""Compressed collectives: the use the reference names for this codec ("compress before a collective",
README.md:72, 103-104) and ships no code for.

    all_gather_compressed(t)       every rank contributes t, every rank receives the concatenation
    all_to_all_compressed(chunks)  rank r sends chunks[d] to rank d and receives one chunk from every rank

Each rank compresses what it sends, the ranks exchange the variable-size archives, every rank
decompresses what it received.  The wire carries about `ratio` x the raw bytes (0.67 for bf16
activations).  The work is cut into `stages` pieces that move through compress -> exchange ->
decompress as a pipeline: while piece k is on the wire (NCCL, asynchronous), piece k+1 is being
compressed and piece k-1 decompressed, and the one host synchronisation an exchange needs (the archive
sizes) waits behind the next piece's codec launch instead of in front of the wire.

The codec work is the ordinary operator path (dietgpu_b200.ops -> C ABI -> sm_100a kernels) on the
caller's device; there is no CPU fallback for it.  The exchange itself is torch.distributed plumbing:
NCCL moves device buffers directly; on a backend without device collectives (gloo, used by the
single-GPU multi-process tests) the packed archives are staged through host memory.

Break-even (profiles/r01_allgather_2gpu.txt): on NVLink 5 (900 GB/s) a plain all-gather of 64 MiB per
rank takes 0.21 ms and the compressed one 0.77 ms, so this only pays on links slower than ~100 GB/s
(inter-node), which is where the reference aims it.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops

_ALIGN = 16  # archives are 16 B aligned in the packed buffer (compressed buffers must be, ans/GpuANSEncode.cu:19-21)


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _device_collectives(group) -> bool:
    return dist.get_backend(group) == "nccl"


def _round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def pack_offsets(sizes: Sequence[int]) -> Tuple[List[int], int]:
    """Offsets of archives of the given byte sizes in one packed buffer, each 16 B aligned; total bytes."""
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += _round_up(int(s), _ALIGN)
    return offs, o


class _Exchange:
    """One in-flight archive exchange (all-gather of every rank's packed archives)."""

    def __init__(self, rows: Sequence[torch.Tensor], group):
        n = len(rows)
        if n == 0:
            raise ValueError("exchange_archives: empty contribution")
        self.dev = rows[0].device
        for r in rows:
            if r.dtype != torch.uint8 or r.dim() != 1 or r.device != self.dev:
                raise ValueError("exchange_archives: archives must be uint8 1-D tensors on one device")
        self.rows, self.group, self.n = list(rows), group, n
        self.world = _world(group)
        self.work = None
        self.sizes: Optional[List[List[int]]] = None
        if self.world == 1:
            return
        self.on_device = _device_collectives(group)
        cdev = self.dev if self.on_device else torch.device("cpu")
        local_sizes = torch.tensor([r.numel() for r in rows], dtype=torch.int32, device=cdev)
        all_sizes = torch.empty(self.world * n, dtype=torch.int32, device=cdev)
        dist.all_gather_into_tensor(all_sizes, local_sizes, group=group)
        self.sizes = all_sizes.cpu().view(self.world, n).tolist()  # the one host sync of the exchange
        totals = [pack_offsets(s)[1] for s in self.sizes]
        self.width = max(max(totals), _ALIGN)
        rank = dist.get_rank(group)
        send = torch.zeros(self.width, dtype=torch.uint8, device=self.dev)
        offs, _ = pack_offsets(self.sizes[rank])
        for r, o in zip(rows, offs):
            send[o:o + r.numel()].copy_(r)
        if self.on_device:
            self.recv = torch.empty(self.world * self.width, dtype=torch.uint8, device=self.dev)
            self.work = dist.all_gather_into_tensor(self.recv, send, group=group, async_op=True)
        else:
            host = torch.empty(self.world * self.width, dtype=torch.uint8)
            dist.all_gather_into_tensor(host, send.cpu(), group=group)
            self.recv = host.to(self.dev)

    def wait(self) -> List[List[torch.Tensor]]:
        if self.world == 1:
            return [self.rows]
        if self.work is not None:
            self.work.wait()
        out: List[List[torch.Tensor]] = []
        for w in range(self.world):
            offs, _ = pack_offsets(self.sizes[w])
            base = w * self.width
            out.append([self.recv[base + o: base + o + s] for o, s in zip(offs, self.sizes[w])])
        return out


def exchange_archives(rows: Sequence[torch.Tensor], group=None) -> List[List[torch.Tensor]]:
    """All ranks contribute the same NUMBER of archives (uint8 1-D tensors of any sizes, one device).
    Returns, for every rank in rank order, the list of that rank's archives (views into one gathered
    buffer on the same device).  Two collectives: sizes, then the padded payload."""
    return _Exchange(rows, group).wait()


def _split(n: int, parts: int, quantum: int) -> List[int]:
    """Boundaries of `parts` nearly equal pieces of [0, n), interior boundaries multiples of `quantum`."""
    parts = max(1, min(parts, max(1, n // max(quantum, 1))))
    per = (n // parts) // quantum * quantum if parts > 1 else n
    return [i * per for i in range(parts)] + [n]


def all_gather_compressed(t: torch.Tensor, group=None, members: int = 8, checksum: bool = False,
                          temp_mem: Optional[torch.Tensor] = None, stages: int = 2) -> torch.Tensor:
    """Every rank contributes the CUDA tensor `t` (same shape and dtype on every rank; fp16 / bf16 / fp32 go
    through the float codec, anything else through the byte codec) and receives the concatenation
    [world * t.numel()] in rank order, bit-exact.  `members` = archives per rank (the codec's parallelism
    comes from blocks, so a handful is enough); `stages` = pipeline pieces (see the module docstring)."""
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("all_gather_compressed: contiguous CUDA tensor expected (no CPU fallback)")
    as_float = t.dtype in (torch.float16, torch.bfloat16, torch.float32)
    flat = t.reshape(-1) if as_float else t.reshape(-1).view(torch.uint8)
    n = flat.numel()
    members = max(1, min(int(members), max(1, n // 4096)))
    # equal member lengths (multiples of 8 elements keep every member 16 B aligned); the last takes the rest
    bounds = _split(n, members, 8)
    members = len(bounds) - 1
    world = _world(group)
    out = torch.empty(world * n, dtype=flat.dtype, device=t.device)
    stages = max(1, min(int(stages), members))
    cuts = [round(k * members / stages) for k in range(stages + 1)]
    pieces = [(cuts[k], cuts[k + 1]) for k in range(stages) if cuts[k + 1] > cuts[k]]

    def compress(a, b):
        comp, sizes, _ = ops.compress_data(as_float, [flat[bounds[i]:bounds[i + 1]] for i in range(a, b)], checksum, temp_mem)
        return comp, sizes

    def decompress(a, b, gathered):
        ins, outs = [], []
        for w in range(world):
            for j, i in enumerate(range(a, b)):
                ins.append(gathered[w][j])
                outs.append(out[w * n + bounds[i]: w * n + bounds[i + 1]])
        # gathered views are 16 B aligned inside the receive buffer; the decoder needs nothing else
        status = torch.zeros(len(ins), dtype=torch.uint8, device=t.device)
        ops.decompress_data(as_float, ins, outs, checksum, temp_mem, status)
        return status

    statuses = []
    nxt = compress(*pieces[0])
    inflight = None  # (piece, exchange)
    for k, (a, b) in enumerate(pieces):
        comp, sizes = nxt
        if k + 1 < len(pieces):
            nxt = compress(*pieces[k + 1])  # queued before the size sync below: runs while the host waits
        hs = sizes.cpu().tolist()
        ex = _Exchange([comp[j, :hs[j]] for j in range(b - a)], group)  # payload goes out asynchronously (NCCL)
        if inflight is not None:
            (pa, pb), pex = inflight
            statuses.append(decompress(pa, pb, pex.wait()))  # previous piece decodes while this one is on the wire
        inflight = ((a, b), ex)
    (pa, pb), pex = inflight
    statuses.append(decompress(pa, pb, pex.wait()))
    if not all(bool(s.all()) for s in statuses):
        raise RuntimeError("all_gather_compressed: a received archive failed to decode")
    return out if as_float else out.view(t.dtype)


def all_to_all_compressed(chunks: Sequence[torch.Tensor], group=None, checksum: bool = False,
                          temp_mem: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """chunks[d] (CUDA, contiguous, same dtype and shape on every rank for a given (source, destination) pair is
    NOT required: sizes travel with the data) goes to rank d; returns [world] tensors, entry s = what rank s
    sent here, bit-exact.  One archive per destination; the payload moves with all_to_all_single and exact
    split sizes, so each link carries only its own compressed chunk."""
    world = _world(group)
    if len(chunks) != world:
        raise ValueError("all_to_all_compressed: one chunk per rank expected")
    dt, dev = chunks[0].dtype, chunks[0].device
    for c in chunks:
        if not c.is_cuda or not c.is_contiguous() or c.dtype != dt or c.device != dev:
            raise ValueError("all_to_all_compressed: contiguous CUDA tensors of one dtype on one device expected")
    as_float = dt in (torch.float16, torch.bfloat16, torch.float32)
    flats = [c.reshape(-1) if as_float else c.reshape(-1).view(torch.uint8) for c in chunks]
    comp, sizes, _ = ops.compress_data(as_float, flats, checksum, temp_mem)
    if world == 1:
        out = torch.empty_like(flats[0])
        ops.decompress_data(as_float, [comp[0, :int(sizes[0])]], [out], checksum, temp_mem)
        return [out.view(chunks[0].shape) if as_float else out.view(dt).view(chunks[0].shape)]
    on_device = _device_collectives(group)
    cdev = dev if on_device else torch.device("cpu")
    # (archive bytes, element count) per destination -> per source
    meta_out = torch.stack([sizes.to(torch.int64), torch.tensor([f.numel() for f in flats], device=dev)], dim=1).to(cdev)
    meta_in = torch.empty_like(meta_out)
    dist.all_to_all_single(meta_in, meta_out, group=group)
    send_sizes = [_round_up(int(v), _ALIGN) for v in meta_out[:, 0].tolist()]  # host sync
    recv_meta = meta_in.tolist()
    recv_sizes = [_round_up(int(v[0]), _ALIGN) for v in recv_meta]
    send = torch.zeros(sum(send_sizes), dtype=torch.uint8, device=dev)
    o = 0
    for d in range(world):
        k = int(meta_out[d, 0])
        send[o:o + k].copy_(comp[d, :k])
        o += send_sizes[d]
    if on_device:
        recv = torch.empty(sum(recv_sizes), dtype=torch.uint8, device=dev)
        dist.all_to_all_single(recv, send, recv_sizes, send_sizes, group=group)
    else:
        host = torch.empty(sum(recv_sizes), dtype=torch.uint8)
        dist.all_to_all_single(host, send.cpu(), recv_sizes, send_sizes, group=group)
        recv = host.to(dev)
    ins, outs, o = [], [], 0
    for s in range(world):
        k, elems = int(recv_meta[s][0]), int(recv_meta[s][1])
        ins.append(recv[o:o + k])
        outs.append(torch.empty(elems, dtype=flats[0].dtype, device=dev))
        o += recv_sizes[s]
    status = torch.zeros(world, dtype=torch.uint8, device=dev)
    ops.decompress_data(as_float, ins, outs, checksum, temp_mem, status)
    if not bool(status.all()):
        raise RuntimeError("all_to_all_compressed: a received archive failed to decode")
    return [x if as_float else x.view(dt) for x in outs]
