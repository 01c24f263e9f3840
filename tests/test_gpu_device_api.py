"""include/dietgpu_b200_device.cuh (warp / CTA level device API; reference README.md:105)
through the example user kernels of tests/cpp/device_api_example.cu: a kernel that compresses the tile
it holds in shared memory must write archives byte-identical to the CPU oracle's, decodable by the
library's own decoder; the device-level decoder must read library-written archives."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import exp_bytes, zipf_bytes
from oracle import oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _lib():
    path = os.path.join(HERE, "cpp", "libdevice_api_example.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: python -c 'import __graft_entry__ as g; g.build()'")
    L = C.CDLL(path)
    L.example_tile_compress.restype = C.c_int
    L.example_tile_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.example_tile_decompress.restype = C.c_int
    L.example_tile_decompress.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    return L


@pytest.mark.parametrize("tile,total", [(32768, 32768 * 5 + 1234), (4096, 4096 * 3), (10000, 70001), (32768, 17)])
def test_user_kernel_archives_equal_oracle(tile, total):
    import dietgpu_b200 as dg

    L = _lib()
    data = np.concatenate([zipf_bytes(total // 2, 1.1, 3), exp_bytes(total - total // 2, 30.0, 4)])
    d_in = torch.from_numpy(data).cuda()
    tiles = (total + tile - 1) // tile
    stride = dg.max_any_compressed_size(tile)
    d_out = torch.zeros((tiles, stride), dtype=torch.uint8, device="cuda")
    d_sz = torch.zeros(tiles, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert L.example_tile_compress(d_in.data_ptr(), total, tile, d_out.data_ptr(), stride, d_sz.data_ptr(), st) == 0
    torch.cuda.synchronize()
    sizes = d_sz.cpu().tolist()
    rows = []
    for i in range(tiles):
        part = data[i * tile:(i + 1) * tile]
        want = O.ans_encode(part, 10)
        got = d_out[i, :sizes[i]].cpu().numpy()
        assert sizes[i] == want.size
        assert np.array_equal(got, want), f"tile {i}: archive differs from the oracle"
        rows.append(d_out[i, :sizes[i]])
    # the library's decoder reads what the user kernel wrote
    outs = [torch.empty(min(tile, total - i * tile), dtype=torch.uint8, device="cuda") for i in range(tiles)]
    dg.decompress_data(False, rows, outs)
    assert np.array_equal(torch.cat(outs).cpu().numpy(), data)
    # and the device-level decoder reads them too
    d_back = torch.zeros(tiles * tile, dtype=torch.uint8, device="cuda")
    d_ok = torch.zeros(tiles, dtype=torch.int32, device="cuda")
    assert L.example_tile_decompress(d_out.data_ptr(), stride, tiles, d_back.data_ptr(), tile, d_ok.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert d_ok.cpu().tolist() == [1] * tiles
    assert np.array_equal(d_back.cpu().numpy()[:total], data)


def test_device_decoder_reads_library_archives():
    import dietgpu_b200 as dg

    L = _lib()
    tile = 32768
    parts = [zipf_bytes(tile, 1.3, i) for i in range(4)] + [exp_bytes(9000, 10.0, 9)]
    ts = [torch.from_numpy(p).cuda() for p in parts]
    comp, sizes, _ = dg.compress_data(False, ts)
    stride = comp.size(1)
    d_back = torch.zeros(len(parts) * tile, dtype=torch.uint8, device="cuda")
    d_ok = torch.zeros(len(parts), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert L.example_tile_decompress(comp.data_ptr(), stride, len(parts), d_back.data_ptr(), tile, d_ok.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert d_ok.cpu().tolist() == [1] * len(parts)
    back = d_back.cpu().numpy()
    for i, p in enumerate(parts):
        assert np.array_equal(back[i * tile:i * tile + p.size], p)
