import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle and the CUDA library exist (cross-compiles without a GPU)."""
    from dietgpu_b200 import capi
    from oracle import oracle as O

    O.build()
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return True


# ---- synthetic inputs shared by the tests (seeded) ------

def exp_bytes(n, lam, seed):
    """bytes ~ min(Exp(lam), 1) * 255, the shape of the reference's generateSymbols
    (ans/ANSTest.cu:18-31)."""
    rng = np.random.default_rng(seed)
    return (np.minimum(rng.exponential(1.0 / lam, n), 1.0) * 255).astype(np.uint8)


def zipf_bytes(n, s, seed):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, 257) ** s
    p /= p.sum()
    return rng.choice(256, size=n, p=p).astype(np.uint8)


def normal_words(n, kind, seed, relu=False):
    """N(0,1) floats as raw words: kind in {'bf16','f16','f32'} (float/FloatTest.cu:109-120)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g, dtype=torch.float32)
    if relu:
        x = torch.relu(x)
    if kind == "bf16":
        return x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    if kind == "f16":
        return x.to(torch.float16).view(torch.int16).numpy().view(np.uint16)
    return x.view(torch.int32).numpy().view(np.uint32)
