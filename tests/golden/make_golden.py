"""Generates tests/golden/oracle_golden.npz: small seeded inputs with the archives the CPU oracle
produces for them.  The oracle is pinned by the reference's own known-answer tests
(ANSStatisticsTest.cu:127-167, checked in tests/test_oracle.py) and by tests/golden/ref_golden.npz,
which holds archives produced by the UNMODIFIED reference on a B200 (make_ref_golden.py).
Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import exp_bytes, normal_words, zipf_bytes  # noqa: E402

from oracle import oracle as O  # noqa: E402

out = {}
cases = {
    "exp20_10013": exp_bytes(10013, 20, 1), "zipf1_8192": zipf_bytes(8192, 1.0, 2), "one": exp_bytes(1, 1, 3),
    "quirk_ids_100_160": np.random.default_rng(4).integers(100, 160, 6000).astype(np.uint8),
    "uniform_5000": np.random.default_rng(5).integers(0, 256, 5000, dtype=np.uint8),
    "single_symbol": np.full(4100, 9, np.uint8), "empty": np.zeros(0, np.uint8),
}
for name, data in cases.items():
    out[f"ans/{name}/in"] = data
    for pb in (9, 10, 11):
        out[f"ans/{name}/pb{pb}"] = O.ans_encode(data, pb, True)
for kind, ft in (("bf16", O.BF16), ("f16", O.F16), ("f32", O.F32)):
    for n in (1, 4099, 20000):
        w = normal_words(n, kind, 100 + n)
        out[f"float/{kind}/{n}/in"] = w
        out[f"float/{kind}/{n}/pb10"] = O.float_compress(ft, w, 10, True)
np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
print("wrote", len(out), "arrays")
