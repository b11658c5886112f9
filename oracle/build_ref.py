"""Build the parts of the REFERENCE that compile from their own few source files into oracle/_ref/ (git-ignored, but it
travels to the GPU box with the snapshot).  Test infrastructure: the product never loads anything from here.

  * efficient_quantile: maua/audiovisual/audioreactive/selfsupervised/features/efficient_quantile/efficient_quantile.cpp
    (one C++ file against libtorch; the reference JIT-builds it through its setup.py).  Compiled here with g++ directly,
    from the source where it lies under /root/reference -- no reference source is copied into this repository.

usage: python oracle/build_ref.py [reference root]      (default /root/reference; a missing reference is not an error:
the GPU box only uses the prebuilt files)
"""
import os
import sys

# run as a script, this directory leads sys.path and oracle/signal.py would shadow the standard library's signal
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.abspath(__file__))]

import subprocess  # noqa: E402
import sysconfig  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
EQ_SRC = "maua/audiovisual/audioreactive/selfsupervised/features/efficient_quantile/efficient_quantile.cpp"
EQ_SO = os.path.join(OUT, "efficient_quantile.so")


def build_efficient_quantile(ref_root):
    import torch
    from torch.utils import cpp_extension as ce

    src = os.path.join(ref_root, EQ_SRC)
    if not os.path.exists(src):
        return None
    if os.path.exists(EQ_SO) and os.path.getmtime(EQ_SO) >= os.path.getmtime(src):
        return EQ_SO
    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    lib_dir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DTORCH_EXTENSION_NAME=efficient_quantile",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *inc, src, "-o", EQ_SO + ".tmp", f"-L{lib_dir}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python",
           f"-Wl,-rpath,{lib_dir}"]
    subprocess.run(cmd, check=True)
    os.replace(EQ_SO + ".tmp", EQ_SO)
    return EQ_SO


def load_efficient_quantile():
    """The reference's compiled ``_efficient_quantile`` (x, q, ignore_nan, method) or None when it was not built."""
    if not os.path.exists(EQ_SO):
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location("efficient_quantile", EQ_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod._efficient_quantile


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    print("efficient_quantile:", build_efficient_quantile(root) or f"reference not found under {root}: skipped")
