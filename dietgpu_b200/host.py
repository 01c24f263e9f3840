"""Host-buffer front end: pinned host tensors in, archives in pinned host memory out, and back.

The reference's operators (dietgpu/DietGpu.cpp:915-937) take device tensors only, so a caller whose
data lives in host memory copies the whole batch up, runs the operator, and copies the result down:
three phases that use the PCIe link in one direction at a time and leave it idle while the kernels
run.  The codec kernels here move >1 TB/s, so for host data the link is the whole cost.  HostCodec
cuts the batch into groups of members (members are independent) and runs
upload / codec call / download of different groups on three CUDA streams: both PCIe directions are
busy at once and the kernels disappear behind the copies.  Archives are byte-for-byte what
compress_data produces for the same members (a member's archive does not depend on its batch).

Everything below is stream plumbing around dietgpu_b200.ops; the codec work is the same C-ABI calls.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class HostCodec:
    """Reusable context for one batch shape: device staging buffers, three streams, pinned size buffers.

    like        : host (or device) tensors giving the member shapes and dtype; uint8 for the byte codec
    groups      : number of member groups the batch is cut into (pipeline depth); clamped to the batch
    """

    def __init__(self, compress_as_float: bool, like: Sequence[torch.Tensor], device=None, groups: int = 8,
                 checksum: bool = False, prob_bits: int = ops.K_DEFAULT_PRECISION):
        if not torch.cuda.is_available():
            raise RuntimeError("dietgpu_b200 has no CPU fallback: HostCodec needs a CUDA device")
        self.as_float, self.checksum, self.prob_bits = bool(compress_as_float), bool(checksum), int(prob_bits)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n = self.n = len(like)
        ops._check(n > 0, "empty batch")
        with torch.cuda.device(self.dev):
            self.dev_in = [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in like]
            self.dev_out = [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in like]
            _, cols = (ops.max_float_compressed_output_size(self.dev_in) if self.as_float
                       else ops.max_any_compressed_output_size(self.dev_in))
            self.cols = cols
            self.comp = torch.empty((n, cols), dtype=torch.uint8, device=self.dev)
            self.sizes = torch.zeros(n, dtype=torch.int32, device=self.dev)
            self.status = torch.zeros(n, dtype=torch.uint8, device=self.dev)
            self.words = torch.zeros(n, dtype=torch.int32, device=self.dev)
            self.host_sizes = torch.zeros(n, dtype=torch.int32).pin_memory()
            self.host_status = torch.zeros(n, dtype=torch.uint8).pin_memory()
            self.up, self.k, self.dn = (torch.cuda.Stream(self.dev) for _ in range(3))
        g = max(1, min(int(groups), n))
        # contiguous groups with (almost) equal byte counts
        weights = [t.numel() * t.element_size() for t in like]
        total, acc, bounds = sum(weights), 0, [0]
        for i, w in enumerate(weights):
            acc += w
            if len(bounds) < g and acc * g >= total * len(bounds) and i + 1 < n:
                bounds.append(i + 1)
        bounds.append(n)
        self.bounds = [(bounds[j], bounds[j + 1]) for j in range(len(bounds) - 1) if bounds[j + 1] > bounds[j]]
        self.temp: Optional[torch.Tensor] = None

    def max_archive_bytes(self) -> int:
        """Row size a host archive matrix [n, cols] needs (reference: max_*_compressed_output_size)."""
        return self.cols

    def _temp_for(self, need: int) -> torch.Tensor:
        if self.temp is None or self.temp.numel() < need:
            self.temp = torch.empty(need + 256, dtype=torch.uint8, device=self.dev)
        return self.temp

    def compress(self, host_in: Sequence[torch.Tensor], host_comp: torch.Tensor) -> List[int]:
        """host_in[i] (pinned) -> archive i in host_comp[i, :size_i] (pinned uint8 [n, >= cols]).
        Returns the archive sizes; host_comp is complete when the call returns."""
        ops._check(len(host_in) == self.n and host_comp.dim() == 2 and host_comp.size(0) >= self.n)
        ops._check(host_comp.dtype == torch.uint8 and host_comp.size(1) >= self.cols)
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            self.up.wait_stream(cur)
            self.k.wait_stream(cur)
            self.dn.wait_stream(cur)
            ev_sz = []
            for a, b in self.bounds:
                with torch.cuda.stream(self.up):
                    for i in range(a, b):
                        self.dev_in[i].copy_(host_in[i], non_blocking=True)
                    ev_up = torch.cuda.Event()
                    ev_up.record(self.up)
                with torch.cuda.stream(self.k):
                    self.k.wait_event(ev_up)
                    if self.temp is None:  # size the scratch once, on the first (largest enough) group
                        _, _, need = ops.compress_data(self.as_float, self.dev_in[a:b], self.checksum, None,
                                                       self.comp[a:b], self.sizes[a:b], prob_bits=self.prob_bits)
                        self._temp_for(2 * need)
                    else:
                        _, _, need = ops.compress_data(self.as_float, self.dev_in[a:b], self.checksum, self.temp,
                                                       self.comp[a:b], self.sizes[a:b], prob_bits=self.prob_bits)
                        self._temp_for(need)
                    ev_k = torch.cuda.Event()
                    ev_k.record(self.k)
                with torch.cuda.stream(self.dn):
                    self.dn.wait_event(ev_k)
                    self.host_sizes[a:b].copy_(self.sizes[a:b], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(self.dn)
                    ev_sz.append(e)
            # the exact-size downloads need the sizes on the host: group by group, while later groups
            # are still uploading / encoding
            for (a, b), e in zip(self.bounds, ev_sz):
                e.synchronize()
                with torch.cuda.stream(self.dn):
                    for i in range(a, b):
                        sz = int(self.host_sizes[i])
                        host_comp[i, :sz].copy_(self.comp[i, :sz], non_blocking=True)
            self.dn.synchronize()
            cur.wait_stream(self.k)
        return [int(v) for v in self.host_sizes.tolist()]

    def decompress(self, host_rows: Sequence[torch.Tensor], host_out: Sequence[torch.Tensor]) -> None:
        """host_rows[i]: pinned uint8 1-D archive i (exact or padded length) -> host_out[i] (pinned).
        Raises RuntimeError if a member fails (capacity / header) or, with checksum=True, on a mismatch."""
        ops._check(len(host_rows) == self.n and len(host_out) == self.n)
        with torch.cuda.device(self.dev):
            cur = torch.cuda.current_stream(self.dev)
            self.up.wait_stream(cur)
            self.k.wait_stream(cur)
            self.dn.wait_stream(cur)
            for a, b in self.bounds:
                rows = []
                with torch.cuda.stream(self.up):
                    for i in range(a, b):
                        sz = host_rows[i].numel()
                        ops._check(host_rows[i].dtype == torch.uint8 and sz <= self.cols)
                        self.comp[i, :sz].copy_(host_rows[i], non_blocking=True)
                        rows.append(self.comp[i, :sz])
                    ev_up = torch.cuda.Event()
                    ev_up.record(self.up)
                with torch.cuda.stream(self.k):
                    self.k.wait_event(ev_up)
                    need = ops.decompress_data(self.as_float, rows, self.dev_out[a:b], self.checksum, self.temp,
                                               self.status[a:b], self.words[a:b], prob_bits=self.prob_bits)
                    self._temp_for(need)
                    ev_k = torch.cuda.Event()
                    ev_k.record(self.k)
                with torch.cuda.stream(self.dn):
                    self.dn.wait_event(ev_k)
                    for i in range(a, b):
                        host_out[i].copy_(self.dev_out[i], non_blocking=True)
            with torch.cuda.stream(self.dn):
                self.host_status.copy_(self.status, non_blocking=True)
            self.dn.synchronize()
            cur.wait_stream(self.k)
        if not bool(self.host_status.all()):
            bad = [i for i, v in enumerate(self.host_status.tolist()) if not v]
            raise RuntimeError(f"HostCodec.decompress: members {bad[:8]} failed (capacity or header)")
