"""Loader of libdietgpu_b200_torch.so: registers torch.ops.dietgpu.* (the reference's ten operators,
same schemas as /root/reference/dietgpu/DietGpu.cpp:915-937) implemented in C++ on the C ABI."""
from __future__ import annotations

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
TORCH_LIB_PATH = os.path.join(_HERE, "libdietgpu_b200_torch.so")
_loaded = False


def load():
    """torch.ops.load_library(...) once; returns torch.ops.dietgpu.  No fallback: raises if missing."""
    global _loaded
    import torch

    if not _loaded:
        if not os.path.exists(TORCH_LIB_PATH):
            raise ImportError(f"{TORCH_LIB_PATH} not found; run `python dietgpu_b200/csrc/build_torch_ops.py`")
        torch.ops.load_library(TORCH_LIB_PATH)
        _loaded = True
    return torch.ops.dietgpu
