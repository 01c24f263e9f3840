"""Host-side mirror of the reference's PyTorch operator surface
(torch.ops.dietgpu.*, /root/reference/dietgpu/DietGpu.cpp:915-972) on top of
the C ABI: same names, same argument meaning, same return values, same error
behaviour (TORCH_CHECK -> RuntimeError).  Tensors are only used for device
memory, streams and dtypes; all compute is in libdietgpu_b200.so.

    compress_data(compress_as_float, ts_in, checksum=False, temp_mem=None,
                  out_compressed=None, out_compressed_bytes=None)
        -> (Tensor [B, maxCols] u8, Tensor [B] i32, int temp_bytes_used)   DietGpu.cpp:277-308
    decompress_data(compress_as_float, ts_in, ts_out, checksum=False, temp_mem=None,
                    out_status=None, out_decompressed_words=None) -> int     DietGpu.cpp:646-677
    ... and the *_split_size / *_simple / max_* variants.

Python's precision is fixed at 10 bits as in the reference (DietGpu.cpp:114);
`prob_bits=` is an extension used by the parity tests (C++ API parity).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import capi

K_DEFAULT_PRECISION = 10  # DietGpu.cpp:114

_FLOAT_TYPES = {torch.float16: capi.FLOAT16, torch.bfloat16: capi.BFLOAT16, torch.float32: capi.FLOAT32}
_DTYPE_OF = {v: k for k, v in _FLOAT_TYPES.items()}


def _check(cond: bool, msg: str = "") -> None:
    if not cond:
        raise RuntimeError("dietgpu_b200 check failed" + (": " + msg if msg else ""))


def _float_type(t: torch.Tensor) -> int:
    _check(t.dtype in _FLOAT_TYPES, f"unsupported float dtype {t.dtype}")
    return _FLOAT_TYPES[t.dtype]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _total_and_max(ts: Sequence[torch.Tensor]) -> Tuple[int, int]:
    total = mx = 0
    for t in ts:
        _check(t.numel() * t.element_size() <= 0xFFFFFFFF, "tensor exceeds 4 GiB - 1")
        total += t.numel()
        mx = max(mx, t.numel())
    return total, mx


class _Temp:
    """Scratch handling: use the caller's temp_mem when it is large enough, else a
    stream-ordered torch allocation (the reference falls back to a synchronous
    cudaMalloc with a warning, utils/StackDeviceMemory.cpp:119-139)."""

    def __init__(self, temp_mem: Optional[torch.Tensor], need: int, device):
        self.need = need
        self.tensor = None
        if temp_mem is not None:
            _check(temp_mem.is_cuda and temp_mem.is_contiguous(), "temp_mem must be a contiguous CUDA tensor")
            base = temp_mem.data_ptr()
            pad = (-base) % 256
            avail = temp_mem.numel() * temp_mem.element_size() - pad
            if avail >= need:
                self.ptr, self.bytes = base + pad, avail
                return
        self.tensor = torch.empty(need + 256, dtype=torch.uint8, device=device)
        base = self.tensor.data_ptr()
        pad = (-base) % 256
        self.ptr, self.bytes = base + pad, need


# ---------------------------------------------------------------- sizes ----

def _bound(v: int, size: int) -> int:
    """The C ABI returns 0 when the worst-case archive does not fit the format's 32-bit sizes (the
    reference CHECK-aborts there, ans/GpuANSEncode.cu:22): raise instead of sizing a buffer from 0."""
    _check(v != 0 or size == 0, f"input of {size} elements is too large for one archive (32-bit format limit)")
    return v


def max_float_compressed_output_size(ts: Sequence[torch.Tensor]) -> Tuple[int, int]:
    _, mx = _total_and_max(ts)
    return len(ts), _bound(capi.lib().dgb_float_max_compressed_size(_float_type(ts[0]), mx), mx)


def max_float_compressed_size(dtype: torch.Tensor, size: int) -> int:
    return _bound(capi.lib().dgb_float_max_compressed_size(_float_type(dtype), size), size)


def max_any_compressed_output_size(ts: Sequence[torch.Tensor]) -> Tuple[int, int]:
    _, mx = _total_and_max(ts)
    return len(ts), _bound(capi.lib().dgb_ans_max_compressed_size(mx * ts[0].element_size()), mx)


def max_any_compressed_size(nbytes: int) -> int:
    return _bound(capi.lib().dgb_ans_max_compressed_size(nbytes), nbytes)


# ------------------------------------------------------------- compress ----

def _validate_out(out_compressed, out_sizes, n: int, cols: int, dev):
    if out_compressed is not None:
        _check(out_compressed.dtype == torch.uint8 and out_compressed.is_cuda
               and out_compressed.is_contiguous() and out_compressed.dim() == 2)
        _check(out_compressed.size(0) >= n and out_compressed.size(1) >= cols)
        _check(out_compressed.device == dev)
        comp = out_compressed
    else:
        comp = torch.empty((n, cols), dtype=torch.uint8, device=dev)
    if out_sizes is not None:
        _check(out_sizes.dtype == torch.int32 and out_sizes.is_cuda and out_sizes.dim() == 1
               and out_sizes.is_contiguous() and out_sizes.size(0) >= n and out_sizes.device == dev)
        sizes = out_sizes
    else:
        sizes = torch.empty((n,), dtype=torch.int32, device=dev)
    return comp, sizes


def compress_data(compress_as_float: bool, ts_in: Sequence[torch.Tensor], checksum: bool = False,
                  temp_mem: Optional[torch.Tensor] = None, out_compressed: Optional[torch.Tensor] = None,
                  out_compressed_bytes: Optional[torch.Tensor] = None, *,
                  prob_bits: int = K_DEFAULT_PRECISION):
    _check(len(ts_in) > 0, "empty batch")
    dev = ts_in[0].device
    for t in ts_in:
        _check(t.is_cuda and t.is_contiguous() and t.device == dev)
        if compress_as_float:
            _check(t.dtype == ts_in[0].dtype)
            _float_type(t)
    n = len(ts_in)
    _, cols = (max_float_compressed_output_size(ts_in) if compress_as_float
               else max_any_compressed_output_size(ts_in))
    comp, sizes = _validate_out(out_compressed, out_compressed_bytes, n, cols, dev)
    row = comp.size(1)
    in_ptrs = capi.ptr_array([t.data_ptr() for t in ts_in])
    out_ptrs = capi.ptr_array([comp.data_ptr() + i * row for i in range(n)])
    L = capi.lib()
    with torch.cuda.device(dev):
        if compress_as_float:
            ft = _float_type(ts_in[0])
            in_sizes = capi.u32_array([t.numel() for t in ts_in])
            need = L.dgb_float_compress_temp_bytes(ft, n, max(t.numel() for t in ts_in))
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_float_compress_pointer(tmp.ptr, tmp.bytes, ft, prob_bits, int(checksum), n,
                                              in_ptrs, in_sizes, out_ptrs, sizes.data_ptr(), _stream())
        else:
            in_sizes = capi.u32_array([t.numel() * t.element_size() for t in ts_in])
            need = L.dgb_ans_encode_temp_bytes(n, max(t.numel() * t.element_size() for t in ts_in))
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_ans_encode_pointer(tmp.ptr, tmp.bytes, prob_bits, int(checksum), n, in_ptrs,
                                          in_sizes, None, out_ptrs, sizes.data_ptr(), _stream())
    capi.check(rc, "compress_data")
    return comp, sizes, tmp.need


def _matrix_to_tensors(n: int, matrix: torch.Tensor, sizes: torch.Tensor) -> List[torch.Tensor]:
    host = sizes[:n].cpu().tolist()  # syncs, like compressedMatrixToTensors (DietGpu.cpp:75-103)
    flat = matrix.view(-1)
    cols = matrix.size(1)
    return [flat.narrow(0, i * cols, host[i]) for i in range(n)]


def compress_data_split_size(compress_as_float: bool, t_in: torch.Tensor, t_in_split_sizes: torch.Tensor,
                             checksum: bool = False, temp_mem: Optional[torch.Tensor] = None,
                             out_compressed: Optional[torch.Tensor] = None,
                             out_compressed_bytes: Optional[torch.Tensor] = None, *,
                             prob_bits: int = K_DEFAULT_PRECISION):
    dev = t_in.device
    _check(t_in.is_cuda and t_in.is_contiguous())
    ft = _float_type(t_in) if compress_as_float else 0
    if not compress_as_float:
        _check(t_in.data_ptr() % 4 == 0, "start pointer is not aligned")
    _check(t_in_split_sizes.is_contiguous() and t_in_split_sizes.device.type == "cpu"
           and t_in_split_sizes.dtype == torch.int32)
    splits = t_in_split_sizes.tolist()
    n = len(splits)
    for i, s in enumerate(splits):
        _check(s > 0)
        if not compress_as_float and i != n - 1:
            _check(s % 4 == 0, "the size of an interior split is not a multiple of 4 bytes")
    mx = max(splits)
    L = capi.lib()
    cols = _bound(L.dgb_float_max_compressed_size(ft, mx) if compress_as_float else L.dgb_ans_max_compressed_size(mx), mx)
    comp, sizes = _validate_out(out_compressed, out_compressed_bytes, n, cols, dev)
    arr = capi.u32_array(splits)
    with torch.cuda.device(dev):
        if compress_as_float:
            need = L.dgb_float_compress_temp_bytes(ft, n, mx)
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_float_compress_split_size(tmp.ptr, tmp.bytes, ft, prob_bits, int(checksum), n,
                                                 t_in.data_ptr(), arr, comp.data_ptr(), comp.size(1),
                                                 sizes.data_ptr(), _stream())
        else:
            need = L.dgb_ans_encode_temp_bytes(n, mx)
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_ans_encode_split_size(tmp.ptr, tmp.bytes, prob_bits, int(checksum), n,
                                             t_in.data_ptr(), arr, None, comp.data_ptr(), comp.size(1),
                                             sizes.data_ptr(), _stream())
    capi.check(rc, "compress_data_split_size")
    return _matrix_to_tensors(n, comp, sizes), sizes, tmp.need


def compress_data_simple(compress_as_float: bool, ts_in: Sequence[torch.Tensor], checksum: bool = False,
                         temp_mem: Optional[int] = 64 * 1024 * 1024) -> List[torch.Tensor]:
    _check(len(ts_in) > 0)
    scratch = None
    if temp_mem and temp_mem > 0:
        scratch = torch.empty(temp_mem, dtype=torch.uint8, device=ts_in[0].device)
    comp, sizes, _ = compress_data(compress_as_float, ts_in, checksum, scratch, None, None)
    host = sizes.cpu().tolist()
    return [comp[i, :host[i]].clone() for i in range(len(ts_in))]


def pull_archives(compress_as_float: bool, ts_src: Sequence[torch.Tensor], ts_dst: Sequence[torch.Tensor],
                  float_dtype: Optional[torch.dtype] = None, out_bytes: Optional[torch.Tensor] = None) -> None:
    """Copies each archive of ts_src (uint8 rows, e.g. views of a peer GPU's memory) into the matching row of
    ts_dst, exactly as long as its header says (dgb_archives_pull): the transfer step of the compressed
    collectives.  `float_dtype` names the float kind of the archives when compress_as_float."""
    _check(len(ts_src) > 0 and len(ts_src) == len(ts_dst))
    dev = ts_dst[0].device
    for a, b in zip(ts_src, ts_dst):
        _check(a.is_cuda and b.is_cuda and b.device == dev and a.dtype == torch.uint8 and b.dtype == torch.uint8)
        _check(a.is_contiguous() and b.is_contiguous())
    ft = _float_type(torch.empty(0, dtype=float_dtype)) if compress_as_float else 0
    if out_bytes is not None:
        _check(out_bytes.is_cuda and out_bytes.dtype == torch.int32 and out_bytes.numel() >= len(ts_src))
    with torch.cuda.device(dev):
        rc = capi.lib().dgb_archives_pull(ft, len(ts_src), capi.ptr_array([t.data_ptr() for t in ts_src]),
                                          capi.ptr_array([t.data_ptr() for t in ts_dst]),
                                          capi.u32_array([t.numel() for t in ts_dst]),
                                          out_bytes.data_ptr() if out_bytes is not None else None, _stream())
    capi.check(rc, "pull_archives")


# ----------------------------------------------------------- decompress ----

def _validate_status(out_status, out_sizes, n, dev):
    if out_status is not None:
        _check(out_status.is_contiguous() and out_status.is_cuda and out_status.dtype == torch.uint8
               and out_status.numel() == n and out_status.device == dev)
    if out_sizes is not None:
        _check(out_sizes.is_contiguous() and out_sizes.is_cuda and out_sizes.dtype == torch.int32
               and out_sizes.numel() == n and out_sizes.device == dev)
    return (out_status.data_ptr() if out_status is not None else None,
            out_sizes.data_ptr() if out_sizes is not None else None)


def _raise_checksum(rc: int, what: str, as_float: bool):
    if rc == capi.ERR_CHECKSUM:
        raise RuntimeError(("floatDecompress" if as_float else "ANSDecode")
                           + ": checksum mismatch seen on decoded data; archive cannot be unpacked")
    capi.check(rc, what)


def decompress_data(compress_as_float: bool, ts_in: Sequence[torch.Tensor], ts_out: Sequence[torch.Tensor],
                    checksum: bool = False, temp_mem: Optional[torch.Tensor] = None,
                    out_status: Optional[torch.Tensor] = None,
                    out_decompressed_words: Optional[torch.Tensor] = None, *,
                    prob_bits: int = K_DEFAULT_PRECISION) -> int:
    _check(len(ts_in) > 0 and len(ts_in) == len(ts_out))
    dev = ts_in[0].device
    n = len(ts_in)
    caps = []
    for ti, to in zip(ts_in, ts_out):
        _check(ti.is_cuda and ti.device == dev and ti.is_contiguous() and ti.dtype == torch.uint8)
        _check(to.is_cuda and to.device == dev and to.is_contiguous())
        if compress_as_float:
            _float_type(to)
        cap = to.numel() if compress_as_float else to.numel() * to.element_size()
        _check(cap <= 0xFFFFFFFF)
        caps.append(cap)
    st_ptr, sz_ptr = _validate_status(out_status, out_decompressed_words, n, dev)
    L = capi.lib()
    in_ptrs = capi.ptr_array([t.data_ptr() for t in ts_in])
    out_ptrs = capi.ptr_array([t.data_ptr() for t in ts_out])
    cap_arr = capi.u32_array(caps)
    with torch.cuda.device(dev):
        if compress_as_float:
            ft = _float_type(ts_out[0])
            need = L.dgb_float_decompress_temp_bytes(ft, n, max(caps))
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_float_decompress_pointer(tmp.ptr, tmp.bytes, ft, prob_bits, int(checksum), n, in_ptrs,
                                                out_ptrs, cap_arr, st_ptr, sz_ptr, None, _stream())
        else:
            need = L.dgb_ans_decode_temp_bytes(n)
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_ans_decode_pointer(tmp.ptr, tmp.bytes, prob_bits, int(checksum), n, in_ptrs, out_ptrs,
                                          cap_arr, st_ptr, sz_ptr, None, _stream())
    _raise_checksum(rc, "decompress_data", compress_as_float)
    return tmp.need


def decompress_data_split_size(compress_as_float: bool, ts_in: Sequence[torch.Tensor], t_out: torch.Tensor,
                               t_out_split_sizes: torch.Tensor, checksum: bool = False,
                               temp_mem: Optional[torch.Tensor] = None,
                               out_status: Optional[torch.Tensor] = None,
                               out_decompressed_words: Optional[torch.Tensor] = None, *,
                               prob_bits: int = K_DEFAULT_PRECISION) -> int:
    _check(len(ts_in) > 0)
    dev = ts_in[0].device
    n = t_out_split_sizes.numel()
    _check(t_out_split_sizes.device.type == "cpu" and t_out_split_sizes.dtype == torch.int32
           and t_out_split_sizes.is_contiguous())
    _check(n == len(ts_in))
    splits = t_out_split_sizes.tolist()
    for ti, s in zip(ts_in, splits):
        _check(ti.is_cuda and ti.device == dev and ti.is_contiguous() and ti.dtype == torch.uint8)
        _check(s > 0)
    _check(t_out.is_cuda and t_out.device == dev and t_out.is_contiguous())
    ft = _float_type(t_out) if compress_as_float else 0
    st_ptr, sz_ptr = _validate_status(out_status, out_decompressed_words, n, dev)
    L = capi.lib()
    in_ptrs = capi.ptr_array([t.data_ptr() for t in ts_in])
    arr = capi.u32_array(splits)
    with torch.cuda.device(dev):
        if compress_as_float:
            need = L.dgb_float_decompress_temp_bytes(ft, n, max(splits))
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_float_decompress_split_size(tmp.ptr, tmp.bytes, ft, prob_bits, int(checksum), n, in_ptrs,
                                                   t_out.data_ptr(), arr, st_ptr, sz_ptr, None, _stream())
        else:
            need = L.dgb_ans_decode_temp_bytes(n)
            tmp = _Temp(temp_mem, need, dev)
            rc = L.dgb_ans_decode_split_size(tmp.ptr, tmp.bytes, prob_bits, int(checksum), n, in_ptrs,
                                             t_out.data_ptr(), arr, st_ptr, sz_ptr, None, _stream())
    _raise_checksum(rc, "decompress_data_split_size", compress_as_float)
    return tmp.need


def decompress_data_simple(compress_as_float: bool, ts_in: Sequence[torch.Tensor], checksum: bool = False,
                           temp_mem: Optional[int] = 64 * 1024 * 1024) -> List[torch.Tensor]:
    _check(len(ts_in) > 0)
    dev = ts_in[0].device
    n = len(ts_in)
    for t in ts_in:
        _check(t.is_cuda and t.device == dev)
    L = capi.lib()
    sizes = torch.empty(n, dtype=torch.int32, device=dev)
    types = torch.zeros(n, dtype=torch.int32, device=dev)
    tmp = _Temp(None, 8 * n + 256, dev)
    in_ptrs = capi.ptr_array([t.data_ptr() for t in ts_in])
    with torch.cuda.device(dev):
        if compress_as_float:
            rc = L.dgb_float_get_compressed_info(tmp.ptr, tmp.bytes, in_ptrs, 0, n, sizes.data_ptr(),
                                                 types.data_ptr(), None, _stream())
        else:
            rc = L.dgb_ans_get_compressed_info(tmp.ptr, tmp.bytes, in_ptrs, 0, n, sizes.data_ptr(), None,
                                               _stream())
    capi.check(rc, "get_compressed_info")
    hs, ht = sizes.cpu().tolist(), types.cpu().tolist()
    outs = []
    for i in range(n):
        if compress_as_float:
            _check(ht[i] == ht[0] and ht[i] in _DTYPE_OF, "inconsistent / invalid float type in archives")
            outs.append(torch.empty(hs[i], dtype=_DTYPE_OF[ht[i]], device=dev))
        else:
            outs.append(torch.empty(hs[i], dtype=torch.uint8, device=dev))
    scratch = None
    if temp_mem and temp_mem >= 256:
        scratch = torch.empty(temp_mem, dtype=torch.uint8, device=dev)
    decompress_data(compress_as_float, ts_in, outs, checksum, scratch, None, None)
    return outs
