"""torch.ops.dietgpu.* (C++ operator library): schema parity on CPU, and on the GPU the flows of the
reference's own Python tests (dietgpu/ans_test.py:50-139, dietgpu/float_test.py:50-178)."""
import numpy as np
import pytest
import torch

from dietgpu_b200 import torch_library

# DietGpu.cpp:915-937, verbatim
REFERENCE_SCHEMAS = {
    "max_float_compressed_output_size": "dietgpu::max_float_compressed_output_size(Tensor[] ts) -> (int, int)",
    "max_float_compressed_size": "dietgpu::max_float_compressed_size(Tensor dtype, int size) -> int",
    "max_any_compressed_output_size": "dietgpu::max_any_compressed_output_size(Tensor[] ts) -> (int, int)",
    "max_any_compressed_size": "dietgpu::max_any_compressed_size(int bytes) -> int",
    "compress_data": "dietgpu::compress_data(bool compress_as_float, Tensor[] ts_in, bool checksum=False, Tensor? temp_mem=None, Tensor? out_compressed=None, Tensor? out_compressed_bytes=None) -> (Tensor, Tensor, int)",
    "compress_data_split_size": "dietgpu::compress_data_split_size(bool compress_as_float, Tensor t_in, Tensor t_in_split_sizes, bool checksum=False, Tensor? temp_mem=None, Tensor? out_compressed=None, Tensor? out_compressed_bytes=None) -> (Tensor[], Tensor, int)",
    "compress_data_simple": "dietgpu::compress_data_simple(bool compress_as_float, Tensor[] ts_in, bool checksum=False, int? temp_mem=67108864) -> Tensor[]",
    "decompress_data": "dietgpu::decompress_data(bool compress_as_float, Tensor[] ts_in, Tensor[] ts_out, bool checksum=False, Tensor? temp_mem=None, Tensor? out_status=None, Tensor? out_decompressed_words=None) -> int",
    "decompress_data_split_size": "dietgpu::decompress_data_split_size(bool compress_as_float, Tensor[] ts_in, Tensor t_out, Tensor t_out_split_sizes, bool checksum=False, Tensor? temp_mem=None, Tensor? out_status=None, Tensor? out_decompressed_words=None) -> int",
    "decompress_data_simple": "dietgpu::decompress_data_simple(bool compress_as_float, Tensor[] ts_in, bool checksum=False, int? temp_mem=67108864) -> Tensor[]",
}


def test_schemas_match_reference():
    ops = torch_library.load()
    for name, want in REFERENCE_SCHEMAS.items():
        got = str(getattr(ops, name).default._schema)
        assert got == want, f"{name}:\n  got  {got}\n  want {want}"
    assert ops.max_any_compressed_size(1 << 20) == 1868320
    assert ops.max_float_compressed_size(torch.empty(0, dtype=torch.bfloat16), 2 << 20) == 16 + 3179040 + 2097152


def _truncated(comp, sizes):
    # ans_test.py:21-26: decode from tensors truncated to exactly the reported size
    return [comp[i, :int(s)].clone() for i, s in enumerate(sizes.cpu())]


@pytest.mark.gpu
def test_ans_flow_like_reference_ans_test():
    ops = torch_library.load()
    dev = torch.device("cuda:0")
    temp = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    # ans_test.py: fp32 N(0,1) tensors compressed bytewise, batched
    ts = [torch.normal(0, 1.0, [n], dtype=torch.float32, device=dev) for n in (10000, 100000, 1000000)]
    rows, cols = ops.max_any_compressed_output_size(ts)
    comp = torch.empty([rows, cols], dtype=torch.uint8, device=dev)
    sizes = torch.zeros([len(ts)], dtype=torch.int, device=dev)
    comp, sizes, used = ops.compress_data(False, ts, False, temp, comp, sizes)
    assert used > 0
    outs = [torch.empty_like(t) for t in ts]
    status = torch.empty([len(ts)], dtype=torch.uint8, device=dev)
    osz = torch.empty([len(ts)], dtype=torch.int32, device=dev)
    ops.decompress_data(False, _truncated(comp, sizes), outs, False, temp, status, osz)
    assert status.cpu().tolist() == [1] * len(ts)
    for a, b in zip(ts, outs):
        assert torch.equal(a, b)
    # empty tensor (ans_test.py empty case), simple API, split-size API
    e = [torch.empty([0], dtype=torch.uint8, device=dev)]
    out = ops.decompress_data_simple(False, ops.compress_data_simple(False, e, True), True)
    assert out[0].numel() == 0
    comp_s = ops.compress_data_simple(False, ts, True)
    for a, b in zip(ts, ops.decompress_data_simple(False, comp_s, True)):
        assert torch.equal(a.view(torch.uint8), b)
    flat = torch.normal(0, 1.0, [4096 * 7], dtype=torch.float32, device=dev).view(torch.uint8)
    splits = torch.tensor([4096 * 4, 4096 * 8, 4096 * 16], dtype=torch.int32)
    rows_l, sz, _ = ops.compress_data_split_size(False, flat, splits, True)
    out = torch.empty_like(flat)
    ops.decompress_data_split_size(False, rows_l, out, splits, True)
    assert torch.equal(out, flat)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16, torch.float32])
def test_float_flow_like_reference_float_test(dt):
    ops = torch_library.load()
    dev = torch.device("cuda:0")
    ts = [torch.normal(0, 1.0, [n], dtype=dt, device=dev) for n in (10000, 100000, 1000000)]
    comp = ops.compress_data_simple(True, ts, True)
    # float_test.py:87-92: must actually shrink
    for t, c in zip(ts, comp):
        assert c.numel() < t.numel() * t.element_size()
    outs = ops.decompress_data_simple(True, comp, True)
    it = torch.int16 if dt != torch.float32 else torch.int32
    for a, b in zip(ts, outs):
        assert b.dtype == dt and torch.equal(a.view(it), b.view(it))
    # batched API with pre-sized outputs, then split-size with and without 16 B alignment
    rows, cols = ops.max_float_compressed_output_size(ts)
    comp_m, sizes, _ = ops.compress_data(True, ts, False)
    assert list(comp_m.shape) == [rows, cols]
    outs = [torch.empty_like(t) for t in ts]
    ops.decompress_data(True, _truncated(comp_m, sizes), outs)
    for a, b in zip(ts, outs):
        assert torch.equal(a.view(it), b.view(it))
    for off in (0, 1):
        flat = torch.normal(0, 1.0, [100000 + off], dtype=dt, device=dev)[off:]
        splits = torch.tensor([1234, 50000, 48766], dtype=torch.int32)
        rows_l, _, _ = ops.compress_data_split_size(True, flat, splits, True)
        out = torch.empty(100000 + off, dtype=dt, device=dev)[off:]
        ops.decompress_data_split_size(True, rows_l, out, splits, True)
        assert torch.equal(out.view(it), flat.view(it))
    # corrupted checksum is reported as in DietGpu.cpp:617-620
    bad = comp[0].clone()
    bad[12] ^= 0x3C
    with pytest.raises(RuntimeError, match="checksum mismatch"):
        ops.decompress_data_simple(True, [bad], True)
