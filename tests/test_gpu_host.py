"""HostCodec (dietgpu_b200/host.py): pinned host tensors -> host archives -> host tensors, pipelined over
member groups on three streams.  Archives must be what the device-side operator produces for the same members
(same sizes as the oracle, decodable by the plain operator), for every group count incl. more groups than members."""
import numpy as np
import pytest
import torch

from conftest import exp_bytes, normal_words, zipf_bytes

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import oracle as O
    return O


def _pin(t):
    return t.pin_memory()


@pytest.mark.parametrize("groups", [1, 3, 8, 64])
def test_host_bytes_roundtrip(groups):
    import dietgpu_b200 as dg
    O = _oracle()
    arrays = [zipf_bytes(200000 + 4099 * i, 1.0, i) for i in range(5)] + [np.zeros(0, np.uint8), exp_bytes(70001, 20, 9)]
    host_in = [_pin(torch.from_numpy(a.copy())) for a in arrays]
    hc = dg.HostCodec(False, host_in, groups=groups, checksum=True)
    host_comp = torch.empty((len(arrays), hc.max_archive_bytes()), dtype=torch.uint8).pin_memory()
    sizes = hc.compress(host_in, host_comp)
    for i, a in enumerate(arrays):
        want = O.ans_encode(a, 10, True)
        assert sizes[i] == want.size
        O.assert_same_ans(host_comp[i, :sizes[i]].numpy(), want, f"member {i}")
    host_out = [_pin(torch.empty_like(t)) for t in host_in]
    hc.decompress([host_comp[i, :sizes[i]] for i in range(len(arrays))], host_out)
    for a, o in zip(arrays, host_out):
        assert np.array_equal(a, o.numpy())


@pytest.mark.parametrize("kind,tdt", [("bf16", torch.bfloat16), ("f16", torch.float16), ("f32", torch.float32)])
def test_host_float_roundtrip(kind, tdt):
    import dietgpu_b200 as dg
    O = _oracle()
    ft = {"bf16": O.BF16, "f16": O.F16, "f32": O.F32}[kind]
    words = [normal_words(30000 + 7777 * i, kind, i) for i in range(9)]
    idt = torch.int32 if kind == "f32" else torch.int16
    host_in = [_pin(torch.from_numpy(w.view(np.int32 if kind == "f32" else np.int16).copy()).view(tdt)) for w in words]
    hc = dg.HostCodec(True, host_in, groups=4)
    host_comp = torch.empty((len(words), hc.max_archive_bytes()), dtype=torch.uint8).pin_memory()
    sizes = hc.compress(host_in, host_comp)
    for i, w in enumerate(words):
        want = O.float_compress(ft, w, 10, False)
        assert sizes[i] == want.size
        O.assert_same_float(host_comp[i, :sizes[i]].numpy(), want, ft, f"member {i}")
    host_out = [_pin(torch.empty_like(t)) for t in host_in]
    hc.decompress([host_comp[i, :sizes[i]] for i in range(len(words))], host_out)
    for t, o in zip(host_in, host_out):
        assert torch.equal(t.view(idt), o.view(idt))
    # a truncated archive must be reported, not decoded into garbage silently
    bad = [host_comp[i, :sizes[i]] for i in range(len(words))]
    bad[2] = host_comp[2, :64].clone().pin_memory()
    bad[2][0:4] = 0
    with pytest.raises(RuntimeError):
        hc.decompress(bad, host_out)


@pytest.mark.parametrize("dt,n", [(torch.bfloat16, 300001), (torch.float32, 70000), (torch.uint8, 123457),
                                  (torch.int32, 50000), (torch.float16, 5)])
def test_all_gather_compressed_one_rank(dt, n):
    # single process (no process group): compress -> pack -> unpack -> decompress must be the identity
    from dietgpu_b200.collectives import all_gather_compressed

    g = torch.Generator(device="cuda").manual_seed(n)
    if dt.is_floating_point:
        t = torch.randn(n, generator=g, device="cuda", dtype=torch.float32).to(dt)
    else:
        t = torch.randint(0, 100, (n,), generator=g, device="cuda", dtype=torch.int32).to(dt)
    out = all_gather_compressed(t, members=5)
    assert out.dtype == dt and out.numel() == n
    assert torch.equal(out.view(torch.uint8), t.view(torch.uint8))
