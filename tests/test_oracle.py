"""CPU tests: the oracle (oracle/dietgpu_oracle.c) against (1) the reference's own known-answer tests,
(2) archives produced by the UNMODIFIED reference on a B200 (tests/golden/ref_golden.npz), (3) its own
round trips on the reference tests' size lists, (4) the committed oracle fixtures."""
import os

import numpy as np
import pytest

from conftest import exp_bytes, normal_words, zipf_bytes
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_normalization_nonzero_kat():
    # ans/ANSStatisticsTest.cu:127-149: 10000 B, every symbol once, the rest symbol 1 -> pdf[1] = 1024 - 255
    h = np.ones(256, np.uint32)
    h[1] = 10000 - 255
    pdf = O.normalize(h, 10000, 10)
    assert pdf[1] == 1024 - 255 and set(np.delete(pdf, 1)) == {1}


def test_normalization_equal_weight_kat():
    # ans/ANSStatisticsTest.cu:151-167: uniform -> every pdf = 4
    pdf = O.normalize(np.full(256, 40, np.uint32), 40 * 256, 10)
    assert set(pdf) == {4}


@pytest.mark.parametrize("pb", [9, 10, 11])
def test_normalization_properties(pb):
    # ans/ANSStatisticsTest.cu:169-207: sum = 2^pb, present => pdf > 0 (absent => 0 unless the
    # symbol-id 'add' quirk applies), within 2x of the true probability for frequent symbols
    rng = np.random.default_rng(pb)
    for lam in (1, 10, 100, 1000):
        d = exp_bytes(12345, lam, int(rng.integers(1 << 30)))
        h = O.histogram(d)
        pdf = O.normalize(h, d.size, pb)
        assert int(pdf.sum()) == 1 << pb
        assert np.all(pdf[h > 0] > 0)
        p = h / d.size
        q = pdf / (1 << pb)
        big = p > 4.0 / (1 << pb)
        assert np.all(q[big] < 2 * p[big]) and np.all(q[big] > 0.5 * p[big])


def test_add_quirk_by_symbol_id():
    # when the first pass is below 2^pb, +1 goes to SYMBOL IDS < diff, present or not
    d = np.random.default_rng(6).integers(100, 160, 30000).astype(np.uint8)
    pdf = O.normalize(O.histogram(d), d.size, 10)
    assert int(pdf.sum()) == 1024
    absent_with_mass = [s for s in range(100) if pdf[s] > 0]
    assert absent_with_mass == list(range(len(absent_with_mass))) and len(absent_with_mass) > 0


def test_size_formulas():
    # SURVEY 8a-12 / 8a-17 values
    assert O.ans_max_compressed_size(1 << 20) == 1868320
    assert O.ans_max_compressed_size(4 << 20) == 5800480
    assert O.float_max_compressed_size(O.BF16, 2 << 20) == 16 + 3179040 + 2097152
    assert O.ans_overhead(0) == 544 and O.ans_overhead(3) == 544 + 3 * 128 + 32


@pytest.mark.parametrize("pb", [9, 10, 11])
def test_ans_roundtrip_reference_size_lists(pb):
    # ans/ANSTest.cu:243-282 size lists; checksum on as there
    for lam in (1.0, 10.0, 100.0, 1000.0):
        for n in (0, 1, 4095, 4096, 4097, 1234, 2345, 3456, 10000, 10013, 8208):
            d = exp_bytes(n, lam, n + int(lam))
            a = O.ans_encode(d, pb, True)
            assert a.size % 16 == 0
            rc, out, got = O.ans_decode(a, pb, verify_checksum=True)
            assert rc == O.OK and got == n and np.array_equal(out, d)


def test_ans_decode_errors():
    d = exp_bytes(5000, 20, 1)
    a = O.ans_encode(d, 10, True)
    rc, _, need = O.ans_decode(a, 10, capacity=4999)
    assert rc == O.ERR_CAPACITY and need == 5000
    assert O.ans_decode(a, 11)[0] == O.ERR_BAD_PROBBITS
    b = a.copy()
    b[3] ^= 0xFF
    assert O.ans_decode(b, 10)[0] == O.ERR_BAD_MAGIC
    c = a.copy()
    c[20] ^= 1
    assert O.ans_decode(c, 10, verify_checksum=True)[0] == O.ERR_CHECKSUM


@pytest.mark.parametrize("kind,ft", [("bf16", O.BF16), ("f16", O.F16), ("f32", O.F32)])
def test_float_roundtrip_and_published_ratio(kind, ft):
    # float/FloatTest.cu sizes + README.md:94 ratios (bf16 ~0.673, fp16 ~0.861)
    for n in (0, 1, 15, 16, 17, 4096, 10000, 65536 + 3):
        w = normal_words(n, kind, 10 + n)
        a = O.float_compress(ft, w, 10, True)
        rc, out, got = O.float_decompress(ft, a, 10, verify_checksum=True)
        assert rc == O.OK and got == n and np.array_equal(out, w)
    w = normal_words(1 << 20, kind, 3)
    ratio = O.float_compress(ft, w, 10).size / w.nbytes
    lo, hi = {"bf16": (0.66, 0.69), "f16": (0.85, 0.87), "f32": (0.82, 0.85)}[kind]
    assert lo < ratio < hi


def test_oracle_matches_reference_archives():
    # archives written by the UNMODIFIED reference (sm_100a build) on a B200: the oracle must produce
    # the same size, pdf, lane states, block table and streams, and decode them bit-exactly
    g = np.load(os.path.join(GOLD, "ref_golden.npz"))
    names = sorted({k.rsplit("/", 1)[0] for k in g.files})
    n_ans = n_float = 0
    for base in names:
        data = g[base + "/in"]
        if base.startswith("ans/"):
            for pb in (9, 10, 11):
                ref = g[f"{base}/pb{pb}"]
                mine = O.ans_encode(data, pb, True)
                assert mine.size == ref.size, (base, pb)
                O.assert_same_ans(mine, ref, f"{base} pb{pb}")
                rc, out, _ = O.ans_decode(ref, pb, verify_checksum=True)
                assert rc == O.OK and np.array_equal(out, data)
                n_ans += 1
        else:
            ft = {"bf16": O.BF16, "f16": O.F16, "f32": O.F32}[base.split("/")[1]]
            ref = g[base + "/pb10"]
            mine = O.float_compress(ft, data, 10, True)
            assert mine.size == ref.size, base
            nc = O.float_noncomp_bytes(ft, data.size)
            assert np.array_equal(mine[:8], ref[:8]) and (mine[8] & 0x1F) == (ref[8] & 0x1F)
            assert np.array_equal(mine[12:16], ref[12:16])  # float-level checksum (first n BYTES)
            if ft == O.F32:
                assert np.array_equal(mine[16:16 + 2 * data.size], ref[16:16 + 2 * data.size])
                o1 = 16 + 2 * ((data.size + 7) // 8 * 8)
                assert np.array_equal(mine[o1:o1 + data.size], ref[o1:o1 + data.size])
            else:
                assert np.array_equal(mine[16:16 + data.size], ref[16:16 + data.size])
            O.assert_same_ans(mine[16 + nc:], ref[16 + nc:], base)
            rc, out, _ = O.float_decompress(ft, ref, 10, verify_checksum=True)
            assert rc == O.OK and np.array_equal(out, data)
            n_float += 1
    assert n_ans >= 18 and n_float >= 9


def test_oracle_golden_fixtures_stable():
    g = np.load(os.path.join(GOLD, "oracle_golden.npz"))
    for k in g.files:
        if k.endswith("/in"):
            continue
        base, tag = k.rsplit("/", 1)
        data, pb = g[base + "/in"], int(tag[2:])
        if base.startswith("ans/"):
            assert np.array_equal(O.ans_encode(data, pb, True), g[k]), k
        else:
            ft = {"bf16": O.BF16, "f16": O.F16, "f32": O.F32}[base.split("/")[1]]
            assert np.array_equal(O.float_compress(ft, data, pb, True), g[k]), k


def test_threaded_oracle_is_deterministic():
    d = zipf_bytes(1 << 20, 1.0, 7)
    O.set_threads(1)
    a = O.ans_encode(d, 10)
    O.set_threads(max(2, O.num_threads()))
    b = O.ans_encode(d, 10)
    assert np.array_equal(a, b)


def test_batch_baseline_equals_per_member_oracle():
    # the whole-batch multi-core entry (bench.py cpu_baseline) writes the same archives, byte for
    # byte, as the per-member restatement, and round-trips
    rng = np.random.default_rng(11)
    byte_batch = [zipf_bytes(70001, 1.0, 1), zipf_bytes(4096, 1.5, 2), np.zeros(0, np.uint8), zipf_bytes(1, 1.0, 3),
                  rng.integers(0, 256, 200000, dtype=np.uint8)]
    for pb in (9, 10, 11):
        archs, outs, _, _ = O.batch_roundtrip(0, byte_batch, pb)
        for a, o, d in zip(archs, outs, byte_batch):
            assert np.array_equal(a, O.ans_encode(d, pb)), pb
            assert np.array_equal(o, d)
    for ft, dt in ((O.BF16, np.uint16), (O.F16, np.uint16), (O.F32, np.uint32)):
        batch = [rng.integers(0, np.iinfo(dt).max, n, dtype=dt) for n in (100003, 4096, 1, 0, 65536)]
        # concentrate the coded byte so that the archives actually compress
        for b in batch:
            if b.size:
                b[::2] = b[0]
        archs, outs, _, _ = O.batch_roundtrip(ft, batch, 10)
        for a, o, d in zip(archs, outs, batch):
            assert np.array_equal(a, O.float_compress(ft, d, 10)), ft
            assert np.array_equal(o, d)


# ---- property tests (hypothesis): any byte string, any precision ----------------------------------------------
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@st.composite
def _byte_arrays(draw):
    n = draw(st.sampled_from([0, 1, 2, 31, 32, 33, 4095, 4096, 4097, 8192, 12289]) | st.integers(0, 20000))
    seed = draw(st.integers(0, 2**31 - 1))
    shape = draw(st.sampled_from(["uniform", "few", "one", "skewed", "two-level"]))
    rng = np.random.default_rng(seed)
    if shape == "uniform":
        a = rng.integers(0, 256, n)
    elif shape == "few":
        a = rng.choice(rng.integers(0, 256, draw(st.integers(1, 6))), n)
    elif shape == "one":
        a = np.full(n, draw(st.integers(0, 255)))
    elif shape == "skewed":
        a = np.minimum(rng.exponential(draw(st.floats(0.5, 60.0)), n), 255)
    else:  # a dominant symbol plus a thin spread over the whole alphabet (pdf-1 symbols, the "add" branch)
        a = np.where(rng.random(n) < 0.97, 7, rng.integers(0, 256, n))
    return a.astype(np.uint8)


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(data=_byte_arrays(), pb=st.sampled_from([9, 10, 11]), cks=st.booleans())
def test_property_ans_roundtrip_any_bytes(data, pb, cks):
    arch = O.ans_encode(data, pb, cks)
    assert arch.size <= O.ans_max_compressed_size(data.size) or data.size == 0
    assert arch.size % 16 == 0
    info = O.ans_info(arch)
    assert info["rc"] == 0 and info["uncompressed"] == data.size and info["prob_bits"] == pb and info["has_checksum"] == cks
    rc, out, got = O.ans_decode(arch, pb, verify_checksum=cks)
    assert rc == 0 and got == data.size and np.array_equal(out, data)
    assert np.array_equal(arch, O.ans_encode(data, pb, cks))  # deterministic
    p = O.parse_ans(arch)
    assert int(np.sum(p["pdf"])) == ((1 << pb) if data.size else 0)
    if data.size:
        # a decoder handed too little room reports it (ans/GpuANSDecode.cuh:326-337) instead of writing past it
        rc2, _, got2 = O.ans_decode(arch, pb, capacity=data.size - 1)
        assert rc2 != 0 and got2 == data.size


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(n=st.integers(0, 9000), seed=st.integers(0, 2**31 - 1), scale=st.sampled_from([1e-3, 1.0, 300.0]),
       kind=st.sampled_from(["bf16", "f16", "f32"]), cks=st.booleans())
def test_property_float_roundtrip(n, seed, scale, kind, cks):
    import torch

    ft = {"bf16": O.BF16, "f16": O.F16, "f32": O.F32}[kind]
    x = torch.randn(n, generator=torch.Generator().manual_seed(seed)) * scale
    if kind == "f32":
        w = x.view(torch.int32).numpy().view(np.uint32)
    else:
        w = x.to(torch.bfloat16 if kind == "bf16" else torch.float16).view(torch.int16).numpy().view(np.uint16)
    arch = O.float_compress(ft, w, 10, cks)
    assert arch.size <= O.float_max_compressed_size(ft, n) or n == 0
    info = O.float_info(arch)
    assert info["rc"] == 0 and info["size"] == n and info["float_type"] == ft
    rc, out, got = O.float_decompress(ft, arch, 10, verify_checksum=cks)
    assert rc == 0 and got == n and np.array_equal(out, w)
