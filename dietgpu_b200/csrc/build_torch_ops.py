"""Builds dietgpu_b200/libdietgpu_b200_torch.so (torch.ops.dietgpu.*) in-tree with g++ against the
installed torch headers; links libdietgpu_b200.so (rpath $ORIGIN).  Called by __graft_entry__.build()."""
import os
import subprocess
import sys

from torch.utils import cpp_extension as ce

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
out = os.path.join(PKG, "libdietgpu_b200_torch.so")
src = os.path.join(HERE, "torch_ops.cpp")
deps = [src, os.path.join(PKG, "..", "include", "dietgpu_b200_compat.hpp"), os.path.join(PKG, "..", "include", "dietgpu_b200.h")]
if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps) and "--force" not in sys.argv:
    sys.exit(0)
import torch  # noqa: E402

inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")]
libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI)),
       *inc, src, "-o", out, f"-L{PKG}", "-ldietgpu_b200", f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda",
       "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{libdir}"]
subprocess.check_call(cmd)
print("built", out)
