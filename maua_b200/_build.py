"""Build libmaua_b200.so in-tree with nvcc for sm_100a (no torch extension machinery, no JIT cache).

The shared library has a plain C ABI (include/maua_b200.h); this module only compiles it.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
# MAUA_B200_LIB: load / build another copy of the library (A/B runs of kernel variants built with extra -D flags)
LIB_PATH = os.environ.get("MAUA_B200_LIB") or os.path.join(LIB_DIR, "libmaua_b200.so")
SOURCES = ["net.cu", "conv_tc.cu", "flrelu.cu", "flrelu_sep.cu", "flrelu_mma.cu", "sg3_misc.cu", "feature_resize.cu", "sg2.cu", "rrdb.cu", "audio.cu", "chroma.cu", "signal_ops.cu", "sequencers.cu", "image_ops.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "maua_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile every CUDA source into maua_b200/lib/libmaua_b200.so (sm_100a, -lineinfo)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmaua_b200.so")
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + _sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force=True, verbose="-v" in sys.argv[1:], extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
