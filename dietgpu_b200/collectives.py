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

Two transports:

  * NCCL (default; any backend torch.distributed offers, gloo for the single-GPU tests): archives are
    packed and exchanged by collectives.  Break-even (profiles/r01_allgather_2gpu.txt): on NVLink 5 a plain
    all-gather of 64 MiB per rank takes 0.21 ms and this one 0.77 ms, so it pays on links slower than
    ~100 GB/s (inter-node), which is where the reference aims it.
  * peer memory (`PeerWorkspace`, one node, NVLink / NVSwitch): every rank encodes into a buffer its peers
    have mapped (torch symmetric memory: CUDA VMM handles exchanged once at set-up); no packing, no size
    exchange (sizes are read from the archive headers on the device), no host synchronisation, CUDA-graph
    capturable.  Three ways to move the bytes, all bit-exact (tests/test_gpu_collectives.py):
      "pull"   one device-side barrier, then the archive mover (dgb_archives_pull: a 64-CTA kernel that copies
               exactly each archive's bytes) reads the peers' archives on a side stream, unit u+1 while the
               decoder works on unit u from local memory;
      "push"   no barrier: as soon as a group of members is coded the mover writes it into every peer's inbox
               and raises a flag there; receivers decode a group when its flag is up;
      "direct" one barrier, then the decode kernel itself reads the peers' archives (its TMA bulk copies and
               cp.async rings take peer addresses like local ones).
    Measured (profiles/r02_allgather_p2p_{2,8}gpu.txt, 256 MiB of bf16 per rank, ratio 0.674): 8 GPUs: plain
    NCCL 3.03 ms, pull 2.87 ms (2.73 replayed from a CUDA graph: 1.11 x plain), push 3.80, direct 3.96;
    2 GPUs: plain NCCL 0.62 ms, push 0.62, pull 0.70, direct 0.84.  The limits: an SM-driven copy over NVLink
    reaches 550-570 GB/s here (the copy engines 728), and the decoder reading peer memory directly is
    latency-bound at 326 GB/s.  A rank encodes once and every peer fetches, so the codec cost is amortised over
    world - 1 links: at 64 MiB per rank the fixed costs still lose (0.72 x plain at 8 GPUs).
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


class PeerWorkspace:
    """A byte buffer of `nbytes` per rank that every rank of `group` (one node) has mapped into its own address
    space.  Setting it up is a collective (allocation + handle exchange); keep it and pass it to every
    `*_compressed(..., peer=ws)` call.  The buffer is used in two halves, alternating per call: the
    barrier of call k+1 orders every rank's reads of call k before any rank's writes of call k+2, so
    one device-side barrier per collective is enough."""

    def __init__(self, nbytes: int, group=None, device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        if dist.get_backend(self.group) != "nccl":
            raise RuntimeError("PeerWorkspace: peer-mapped memory needs one CUDA device per rank (NCCL group)")
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.half = _round_up(int(nbytes), 256)
        self.buf = symm.empty(2 * self.half, dtype=torch.uint8, device=dev)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.calls = 0
        self.timeout_ms = 20000
        self._views = {}
        self._staging = None
        self.side = torch.cuda.Stream(device=dev)  # archive mover: runs beside the decoder

    def begin(self, nbytes: int) -> int:
        """Byte offset of the half this call uses."""
        if nbytes > self.half:
            raise ValueError(f"PeerWorkspace: call needs {nbytes} B per rank, workspace halves hold {self.half} B")
        off = (self.calls & 1) * self.half
        self.calls += 1
        return off

    def view(self, rank: int, offset: int, nbytes: int) -> torch.Tensor:
        """uint8 tensor on THIS rank's device that aliases bytes [offset, offset + nbytes) of `rank`'s buffer."""
        if rank == self.rank:
            return self.buf[offset:offset + nbytes]
        key = (rank, offset, nbytes)
        v = self._views.get(key)
        if v is None:
            if len(self._views) > 512:  # views are cheap to rebuild; do not grow without bound when sizes vary
                self._views.clear()
            v = self._views[key] = self.hdl.get_buffer(rank, (nbytes,), torch.uint8, offset)
        return v

    def staging(self, nbytes: int) -> torch.Tensor:
        """Local (not peer-mapped) buffer the archive mover fills; grown on demand, kept across calls."""
        if self._staging is None or self._staging.numel() < nbytes:
            self._staging = torch.empty(nbytes, dtype=torch.uint8, device=self.buf.device)
        return self._staging[:nbytes]

    def barrier(self) -> None:
        """Device-side barrier of all ranks on the current stream (no host synchronisation)."""
        self.hdl.barrier(channel=0)

    # point-to-point flags on the current stream (channel 0 belongs to barrier()); a lost signal traps the
    # waiting kernel after `timeout_ms` instead of hanging the device
    def signal(self, dst: int, channel: int) -> None:
        self.hdl.put_signal(dst, channel=channel, timeout_ms=self.timeout_ms)

    def wait(self, src: int, channel: int) -> None:
        self.hdl.wait_signal(src, channel=channel, timeout_ms=self.timeout_ms)


def _archive_cols(as_float: bool, dtype: torch.dtype, longest: int) -> int:
    """Row pitch for archives of members of at most `longest` elements (bytes for the byte codec)."""
    b = ops.max_float_compressed_size(torch.empty(0, dtype=dtype), longest) if as_float else ops.max_any_compressed_size(longest)
    return _round_up(int(b), _ALIGN)


def _pull_and_decode(ws: "PeerWorkspace", as_float: bool, dtype: torch.dtype, units, cols: int, checksum: bool,
                     temp_mem: Optional[torch.Tensor], check: bool, direct: bool) -> None:
    """units = [(source rows [k, cols] in peer memory, [k output tensors]), ...].  direct: the decode kernel reads
    the peer rows itself.  Otherwise a two-deep pipeline: the archive mover (dgb_archives_pull, a small grid on
    the workspace's side stream) brings unit u+1 into local staging while the decoder works on unit u."""
    dev = units[0][1][0].device
    statuses = []
    if direct:
        ins = [rows[i] for rows, outs in units for i in range(len(outs))]
        outs = [o for _, os_ in units for o in os_]
        st = torch.zeros(len(ins), dtype=torch.uint8, device=dev)
        ops.decompress_data(as_float, ins, outs, checksum, temp_mem, st)
        statuses.append(st)
    else:
        cur = torch.cuda.current_stream(dev)
        kmax = max(len(outs) for _, outs in units)
        stage = ws.staging(2 * kmax * cols).view(2, kmax, cols)
        start = torch.cuda.Event()
        start.record(cur)
        ws.side.wait_event(start)
        decoded = [None, None]  # event: the decoder is done with staging buffer b
        for u, (rows, outs) in enumerate(units):
            b, k = u & 1, len(outs)
            with torch.cuda.stream(ws.side):
                if decoded[b] is not None:
                    ws.side.wait_event(decoded[b])
                ops.pull_archives(as_float, [rows[i] for i in range(k)], [stage[b, i] for i in range(k)], dtype)
                pulled = torch.cuda.Event()
                pulled.record(ws.side)
            cur.wait_event(pulled)
            st = torch.zeros(k, dtype=torch.uint8, device=dev)
            ops.decompress_data(as_float, [stage[b, i] for i in range(k)], outs, checksum, temp_mem, st)
            statuses.append(st)
            decoded[b] = torch.cuda.Event()
            decoded[b].record(cur)
    if check and not all(bool(st.all()) for st in statuses):
        raise RuntimeError("compressed collective: a peer archive failed to decode")


def _all_gather_push(flat: torch.Tensor, as_float: bool, bounds: List[int], ws: "PeerWorkspace", checksum: bool,
                     temp_mem: Optional[torch.Tensor], check: bool, stages: int) -> torch.Tensor:
    """Push transport: the members are coded in `stages` groups; as soon as a group is coded, the archive mover
    (side stream) WRITES its archives into every peer's inbox -- stores over NVLink are posted, they do not pay the
    round trip a pull does -- and raises a flag there; the receiver decodes a group from its own memory when the
    flag is up.  Coding of group g+1, the pushes of group g and the decoding of what has arrived overlap; there is
    no barrier: the flags of call k+1 order every rank's reads of call k before the writes of call k+2."""
    n, members, world, rank = flat.numel(), len(bounds) - 1, ws.world, ws.rank
    dev = flat.device
    cols = _archive_cols(as_float, flat.dtype, max(bounds[i + 1] - bounds[i] for i in range(members)))
    region = members * cols
    off = ws.begin((world + 1) * region)  # [own archives | inbox of source 0 | ... | inbox of source world-1]
    out = torch.empty(world * n, dtype=flat.dtype, device=dev)
    mine = ws.view(rank, off, region).view(members, cols)
    stages = max(1, min(int(stages), members, 7))
    cuts = [round(k * members / stages) for k in range(stages + 1)]
    groups = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    cur = torch.cuda.current_stream(dev)
    peers = [(rank + k) % world for k in range(1, world)]
    sizes = torch.empty(members, dtype=torch.int32, device=dev)
    for g, (a, b) in enumerate(groups):
        ops.compress_data(as_float, [flat[bounds[i]:bounds[i + 1]] for i in range(a, b)], checksum, temp_mem, mine[a:b], sizes[a:b])
        coded = torch.cuda.Event()
        coded.record(cur)
        with torch.cuda.stream(ws.side):
            ws.side.wait_event(coded)
            for p in peers:
                inbox = ws.view(p, off + (1 + rank) * region, region).view(members, cols)
                ops.pull_archives(as_float, [mine[i] for i in range(a, b)], [inbox[i] for i in range(a, b)], flat.dtype)
                ws.signal(p, 1 + g)
    out[rank * n:(rank + 1) * n].copy_(flat)  # own shard: a local copy, no codec
    statuses = []
    for g, (a, b) in enumerate(groups):
        for src in reversed(peers):  # rank - 1 pushed to this rank first
            ws.wait(src, 1 + g)
            rows = ws.view(rank, off + (1 + src) * region, region).view(members, cols)
            st = torch.zeros(b - a, dtype=torch.uint8, device=dev)
            ops.decompress_data(as_float, [rows[i] for i in range(a, b)],
                                [out[src * n + bounds[i]: src * n + bounds[i + 1]] for i in range(a, b)], checksum, temp_mem, st)
            statuses.append(st)
    done = torch.cuda.Event()
    with torch.cuda.stream(ws.side):
        done.record(ws.side)
    cur.wait_event(done)  # the side stream joins (stream capture needs it; the next call's scratch reuse too)
    if check and not all(bool(st.all()) for st in statuses):
        raise RuntimeError("all_gather_compressed: a peer archive failed to decode")
    return out


def _all_gather_peer(flat: torch.Tensor, as_float: bool, bounds: List[int], ws: "PeerWorkspace", checksum: bool,
                     temp_mem: Optional[torch.Tensor], check: bool, direct: bool, stages: int) -> torch.Tensor:
    n, members, world, rank = flat.numel(), len(bounds) - 1, ws.world, ws.rank
    cols = _archive_cols(as_float, flat.dtype, max(bounds[i + 1] - bounds[i] for i in range(members)))
    off = ws.begin(members * cols)
    out = torch.empty(world * n, dtype=flat.dtype, device=flat.device)
    mine = ws.view(rank, off, members * cols).view(members, cols)
    sizes = torch.empty(members, dtype=torch.int32, device=flat.device)
    ops.compress_data(as_float, [flat[bounds[i]:bounds[i + 1]] for i in range(members)], checksum, temp_mem, mine, sizes)
    ws.barrier()  # every rank's archives are complete (and every rank is done with the previous call's other half)
    out[rank * n:(rank + 1) * n].copy_(flat)  # own shard: a local copy, no codec
    stages = max(1, min(int(stages), members))
    cuts = [round(k * members / stages) for k in range(stages + 1)]
    units = []
    for k in range(1, world):
        w = (rank + k) % world  # start with a different peer on every rank: spreads the pulls over the links
        rows = ws.view(w, off, members * cols).view(members, cols)
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b > a:
                units.append((rows[a:b], [out[w * n + bounds[i]: w * n + bounds[i + 1]] for i in range(a, b)]))
    if units:
        _pull_and_decode(ws, as_float, flat.dtype, units, cols, checksum, temp_mem, check, direct)
    return out


def _split(n: int, parts: int, quantum: int) -> List[int]:
    """Boundaries of `parts` nearly equal pieces of [0, n), interior boundaries multiples of `quantum`."""
    parts = max(1, min(parts, max(1, n // max(quantum, 1))))
    per = (n // parts) // quantum * quantum if parts > 1 else n
    return [i * per for i in range(parts)] + [n]


def all_gather_compressed(t: torch.Tensor, group=None, members: int = 8, checksum: bool = False,
                          temp_mem: Optional[torch.Tensor] = None, stages: int = 2,
                          peer: Optional[PeerWorkspace] = None, check: bool = True,
                          peer_mode: str = "auto") -> torch.Tensor:
    """Every rank contributes the CUDA tensor `t` (same shape and dtype on every rank; fp16 / bf16 / fp32 go
    through the float codec, anything else through the byte codec) and receives the concatenation
    [world * t.numel()] in rank order, bit-exact.  `members` = archives per rank (the codec's parallelism
    comes from blocks, so a handful is enough); `stages` = pipeline pieces (see the module docstring).
    With `peer` (a PeerWorkspace of the same group) the archives move through peer-mapped memory instead of a
    collective; `peer_mode` = "auto" (push for two ranks, pull beyond), "pull" (after a barrier the archive mover reads the peers' archives, unit u+1 while
    unit u is decoded; the fastest at 8 GPUs), "push" (the mover writes each coded group into the peers' inboxes
    and flags it; no barrier; the fastest at 2 GPUs) or "direct" (after a barrier the decode kernel reads peer
    memory itself); `stages` = pipeline groups per peer;
    `check=False` then skips the only host synchronisation (the read of the decode status)."""
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("all_gather_compressed: contiguous CUDA tensor expected (no CPU fallback)")
    as_float = t.dtype in (torch.float16, torch.bfloat16, torch.float32)
    flat = t.reshape(-1) if as_float else t.reshape(-1).view(torch.uint8)
    n = flat.numel()
    members = max(1, min(int(members), max(1, n // 4096)))
    # equal member lengths (multiples of 8 elements keep every member 16 B aligned); the last takes the rest
    bounds = _split(n, members, 8)
    members = len(bounds) - 1
    if peer is not None:
        if peer_mode == "auto":  # measured: push wins with one peer, the pipelined pull with seven
            peer_mode = "push" if peer.world <= 2 else "pull"
        if peer_mode not in ("push", "pull", "direct"):
            raise ValueError("all_gather_compressed: peer_mode is 'auto', 'push', 'pull' or 'direct'")
        if peer_mode == "push":
            out = _all_gather_push(flat, as_float, bounds, peer, checksum, temp_mem, check, stages)
        else:
            out = _all_gather_peer(flat, as_float, bounds, peer, checksum, temp_mem, check, peer_mode == "direct", stages)
        return out if as_float else out.view(t.dtype)
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


def _all_to_all_peer(flats: List[torch.Tensor], as_float: bool, ws: PeerWorkspace, checksum: bool,
                     temp_mem: Optional[torch.Tensor], check: bool, direct: bool) -> List[torch.Tensor]:
    """Regular all-to-all (chunk (s, d) has the same length on every rank pair): rank r encodes its chunks as one
    batch into its own peer-mapped rows, and after the barrier decodes row [r] of every peer straight out of
    the peer's memory."""
    world, rank, dev = ws.world, ws.rank, flats[0].device
    m = flats[0].numel()
    if any(f.numel() != m for f in flats):
        raise ValueError("all_to_all_compressed(peer=...): equal chunk lengths expected (sizes are not exchanged)")
    cols = _archive_cols(as_float, flats[0].dtype, m)
    off = ws.begin(world * cols)
    mine = ws.view(rank, off, world * cols).view(world, cols)
    sizes = torch.empty(world, dtype=torch.int32, device=dev)
    ops.compress_data(as_float, flats, checksum, temp_mem, mine, sizes)
    ws.barrier()
    outs = [torch.empty(m, dtype=flats[0].dtype, device=dev) for _ in range(world)]
    outs[rank].copy_(flats[rank])
    units = []
    for k in range(1, world):
        src = (rank + k) % world
        units.append((ws.view(src, off + rank * cols, cols).view(1, cols), [outs[src]]))
    if units:
        _pull_and_decode(ws, as_float, flats[0].dtype, units, cols, checksum, temp_mem, check, direct)
    return outs


def all_to_all_compressed(chunks: Sequence[torch.Tensor], group=None, checksum: bool = False,
                          temp_mem: Optional[torch.Tensor] = None, peer: Optional[PeerWorkspace] = None,
                          check: bool = True, peer_mode: str = "pull") -> List[torch.Tensor]:
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
    if peer is not None:
        outs = _all_to_all_peer(flats, as_float, peer, checksum, temp_mem, check, peer_mode == "direct")
        return [x if as_float else x.view(dt) for x in outs]
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
