"""Build realtime/ddsp_host (the libtorch C++ consumer of the exported model) in-tree with g++."""
import os
import subprocess
import sys

import torch
from torch.utils import cpp_extension as ce

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ddsp_host")


def build(force=False):
    src = os.path.join(HERE, "ddsp_host.cpp")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > os.path.getmtime(src):
        return OUT
    try:
        paths = ce.include_paths("cuda")
    except TypeError:
        paths = ce.include_paths(cuda=True)
    inc = []
    for p in paths:
        inc += ["-isystem", p]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O2", "-std=c++17", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, "-isystem", "/usr/local/cuda/include",
           src, "-o", OUT, "-L", tlib, "-Wl,--no-as-needed", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10",
           "-lc10_cuda", "-ldl", "-lpthread", f"-Wl,-rpath,{tlib}"]
    print("[realtime.build_host]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build("--force" in sys.argv)
