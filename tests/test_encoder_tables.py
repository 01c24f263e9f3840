"""Host-side statement of the encoder symbol-table arithmetic (dietgpu_b200/csrc/encode.cu normalizeAndPublish,
EncSym, encodeUpdate; common.cuh EncEntry / EncEntryWide), checked exhaustively over pdf and at the edges of
the state range.  The reference derives the quotient with a round-up magic plus an add
(ans/GpuANSStatistics.cuh:343-358, ans/GpuANSEncode.cuh:79-86); the kernels use an add-free reciprocal,
fold a correction for pdf == 1 into the cdf term, and (packed format) rebuild the threshold from 2^pb - pdf.
Each of those steps must leave the state update bit-identical to the reference formula
    x' = (x / pdf) << pb  +  x % pdf  +  cdf            for every 2^15 <= x < 2^31 the coder can hold."""
import numpy as np
import pytest

U64 = np.uint64


def entry_constants(pdf: int, cdf: int, pb: int):
    """(magic, shift, kmp, cdf_term, thr) exactly as normalizeAndPublish builds them."""
    K = 1 << pb
    shift, magic, cdf_term = 0, 0, cdf
    if pdf > 1:
        shift = (pdf - 1).bit_length() - 1          # 31 - clz(pdf - 1)
        magic = ((1 << (32 + shift)) + pdf - 1) // pdf
    elif pdf == 1:
        magic = 0xFFFFFFFF
        cdf_term = cdf + (K - 1)
    return magic, shift, K - pdf, cdf_term, (pdf << (31 - pb)) & 0xFFFFFFFF


def update(x: np.ndarray, magic: int, shift: int, kmp: int, cdf_term: int) -> np.ndarray:
    """encodeUpdate: div = hi32(x * magic) >> shift; x' = div * kmp + x + cdf_term (mod 2^32)."""
    div = ((x.astype(U64) * U64(magic)) >> U64(32)) >> U64(shift)
    return ((div * U64(kmp) + x.astype(U64) + U64(cdf_term)) & U64(0xFFFFFFFF)).astype(np.uint32)


def states_for(pdf: int, pb: int, rng) -> np.ndarray:
    """States the update can see for this symbol: after renormalisation x < pdf << (31 - pb), and x >= 2^15
    unless it was just shifted down (then x >= (pdf << (31 - pb)) >> 16)."""
    thr = pdf << (31 - pb)
    lo = min(1 << 15, thr >> 16)
    edge = [lo, lo + 1, thr - 1, thr - 2, (1 << 15), (1 << 15) + 1, pdf * (thr // pdf) - 1, pdf * ((thr // pdf) - 1)]
    edge = [e for e in edge if lo <= e < thr]
    rnd = rng.integers(lo, thr, 64, dtype=np.int64)
    return np.unique(np.array(edge + rnd.tolist(), dtype=np.int64)).astype(np.uint32)


@pytest.mark.parametrize("pb", [9, 10, 11])
def test_update_equals_reference_formula(pb):
    rng = np.random.default_rng(pb)
    K = 1 << pb
    for pdf in range(1, K + 1):
        cdf = int(rng.integers(0, K - pdf + 1))
        magic, shift, kmp, cdf_term, _ = entry_constants(pdf, cdf, pb)
        assert magic < (1 << 32) and shift < 32
        x = states_for(pdf, pb, rng)
        want = ((x.astype(U64) // U64(pdf)) << U64(pb)) + (x.astype(U64) % U64(pdf)) + U64(cdf)
        got = update(x, magic, shift, kmp, cdf_term)
        assert np.array_equal(got.astype(U64), want), (pb, pdf)
        assert int(want.max()) < (1 << 31) + K          # the coder state stays a 31-bit quantity


@pytest.mark.parametrize("pb", [9, 10, 11])
def test_reciprocal_is_exact_below_2_pow_31(pb):
    # the quotient itself (pdf >= 2) is exact on the whole range [0, 2^31), not only below the threshold
    rng = np.random.default_rng(100 + pb)
    for pdf in range(2, (1 << pb) + 1):
        magic, shift, *_ = entry_constants(pdf, 0, pb)
        top = (1 << 31) - 1
        x = np.array([0, 1, pdf - 1, pdf, top, top - 1, pdf * (top // pdf), pdf * (top // pdf) - 1]
                     + rng.integers(0, 1 << 31, 32).tolist(), dtype=U64)
        assert np.array_equal(((x * U64(magic)) >> U64(32)) >> U64(shift), x // U64(pdf)), pdf


@pytest.mark.parametrize("pb", [9, 10, 11])
def test_packed_entry_fields(pb):
    # EncEntry.pack = shift (bits 0..4) | 2^pb - pdf (bits 5..16) | cdf term (bits 20..31); the loader takes
    # the shift from the low 5 bits (shf.wrap), kmp = (pack >> 5) & 0xfff, cdf = pack >> 20 and rebuilds
    # thr = kmp * -(2^(31-pb)) + 2^31 (mod 2^32)
    K = 1 << pb
    for pdf in range(0, K + 1):
        cdf = K - pdf
        magic, shift, kmp, cdf_term, thr = entry_constants(pdf, cdf, pb) if pdf else (0, 0, K, cdf, 0)
        assert shift < 32 and kmp <= 0xFFF and cdf_term <= 0xFFF
        pack = shift | (kmp << 5) | (cdf_term << 20)
        assert pack < (1 << 32)
        assert (pack & 31) == shift and ((pack >> 5) & 0xFFF) == kmp and (pack >> 20) == cdf_term
        neg_scale = (-(1 << (31 - pb))) & 0xFFFFFFFF
        assert (kmp * neg_scale + 0x80000000) & 0xFFFFFFFF == thr
        # wide entry: cdfShift = shift | cdf term << 5
        wide = shift | (cdf_term << 5)
        assert wide < (1 << 32) and (wide & 31) == shift and (wide >> 5) == cdf_term
