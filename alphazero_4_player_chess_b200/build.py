"""Build the CUDA library in-tree: csrc/*.cu -> libfpc.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfpc.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["fpc_kernels.cu", "fpc_puct.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-cudart", "shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fpc.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    subprocess.check_call(cmd)
    return LIB


BINDING_SRC = os.path.join(CSRC, "binding.cpp")
DROPIN = os.path.join(HERE, "dropin")


def binding_path() -> str:
    import sysconfig
    return os.path.join(DROPIN, "alphazero_cpp" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_binding(force: bool = False) -> str:
    """The pybind11 drop-in module `alphazero_cpp` (csrc/binding.cpp) -> dropin/alphazero_cpp*.so, linked
    against libfpc.so.  `sys.path.insert(0, <package>/dropin)` makes `import alphazero_cpp` resolve to it."""
    import sysconfig

    import pybind11
    build()
    out = binding_path()
    deps = [BINDING_SRC, os.path.join(HERE, "..", "include", "fpc.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(DROPIN, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-Wall",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], BINDING_SRC, "-o", out,
           "-L" + HERE, "-lfpc", "-Wl,-rpath,$ORIGIN/.."]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_binding(force="--force" in sys.argv))
