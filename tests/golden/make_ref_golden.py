"""Runs on the GPU box (gpurun): archives produced by the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/build_ref.sh) for small seeded inputs -> gpurun_out/ref_golden.npz, which is
then committed as tests/golden/ref_golden.npz.  CPU tests compare the oracle with it field by field.
    gpurun -- python tests/golden/make_ref_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import exp_bytes, normal_words, zipf_bytes  # noqa: E402

from oracle import ref_lib  # noqa: E402

ref = ref_lib.RefCodec(256 << 20)
L = ref_lib.lib()
out = {}
cases = {
    "exp20_10013": exp_bytes(10013, 20, 1), "zipf1_70000": zipf_bytes(70000, 1.0, 2), "one": exp_bytes(1, 1, 3),
    "quirk_ids_100_160": np.random.default_rng(4).integers(100, 160, 30000).astype(np.uint8),
    "uniform_50000": np.random.default_rng(5).integers(0, 256, 50000, dtype=np.uint8),
    "single_symbol": np.full(10000, 9, np.uint8),
}
for name, data in cases.items():
    t = torch.from_numpy(data).cuda()
    out[f"ans/{name}/in"] = data
    for pb in (9, 10, 11):
        comp = torch.zeros((1, L.ref_ans_max_compressed_size(data.size)), dtype=torch.uint8, device="cuda")
        sz = torch.zeros(1, dtype=torch.int32, device="cuda")
        ref.ans_encode([t], comp, sz, pb, True)
        torch.cuda.synchronize()
        out[f"ans/{name}/pb{pb}"] = comp[0, :int(sz[0])].cpu().numpy()
dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}
for kind, ft in (("bf16", 2), ("f16", 1), ("f32", 3)):
    for n in (1, 4099, 70001):
        w = normal_words(n, kind, 100 + n)
        t = torch.from_numpy(w.view(np.int16 if kind != "f32" else np.int32).copy()).view(dt[kind]).cuda()
        comp = torch.zeros((1, L.ref_float_max_compressed_size(ft, n)), dtype=torch.uint8, device="cuda")
        sz = torch.zeros(1, dtype=torch.int32, device="cuda")
        ref.float_compress(ft, [t], comp, sz, 10, True)
        torch.cuda.synchronize()
        out[f"float/{kind}/{n}/in"] = w
        out[f"float/{kind}/{n}/pb10"] = comp[0, :int(sz[0])].cpu().numpy()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_golden.npz"), **out)
print("wrote", len(out), "arrays")
