"""GPU test of the C++ drop-in layer: tests/cpp/compat_roundtrip.cu replays the reference's gtest cases
(ANSTest.*, FloatTest.*) through include/dietgpu_b200_compat.hpp, i.e. the reference's own C++ signatures."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_compat_roundtrip():
    exe = os.path.join(ROOT, "tests", "cpp", "compat_roundtrip")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "dietgpu_b200", "csrc"), "-s", "compat_test"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all passed" in r.stdout
