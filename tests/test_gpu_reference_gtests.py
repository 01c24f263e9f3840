"""The reference's OWN googletest suites (dietgpu/ans/ANSTest.cu:243-282, dietgpu/float/FloatTest.cu:
270-311), compiled unchanged from /root/reference against libdietgpu_b200.so through the include-path
shim tests/ref_shim -> include/dietgpu_b200_compat.hpp (tests/cpp/build_ref_tests.sh, run by
__graft_entry__.build() in the container that has the reference; the binaries travel to the GPU box).
SURVEY.md section 8f-1's acceptance line: ANSTest and FloatTest pass against the new library."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(name, expect_tests):
    exe = os.path.join(HERE, "cpp", name)
    if not os.path.exists(exe):
        pytest.fail(f"{exe} missing: run tests/cpp/build_ref_tests.sh where /root/reference exists")
    p = subprocess.run([exe, "--gtest_color=no"], capture_output=True, text=True, timeout=900)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-4000:]
    assert "[  PASSED  ]" in out and "FAILED" not in out, out[-4000:]
    for t in expect_tests:
        assert f"[       OK ] {t}" in out, (t, out[-4000:])


@pytest.mark.gpu
def test_reference_ans_gtest_unchanged():
    _run("ref_ans_test", ["ANSTest.ZeroSized", "ANSTest.BatchPointer", "ANSTest.BatchPointerLarge", "ANSTest.BatchStride"])


@pytest.mark.gpu
def test_reference_float_gtest_unchanged():
    _run("ref_float_test", ["FloatTest.Batch", "FloatTest.LargeBatch", "FloatTest.BatchSize1"])
