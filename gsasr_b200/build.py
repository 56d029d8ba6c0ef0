"""Build recipe for libgsraster.so (the C-ABI library with every CUDA kernel of the path).

Plain ``nvcc -shared`` for sm_100a, in-tree, no torch headers: the library only needs the CUDA
runtime (linked statically), so it loads into any process -- PyTorch, a C++ host, ctypes.

    python -m gsasr_b200.build            # build if stale
    python -m gsasr_b200.build --force
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgsraster.so")

SOURCES = ["gsraster.cu", "gsr_head.cu"]
HEADERS = [
    "gsr_common.cuh",
    "gsr_prepass.cuh",
    "gsr_forward.cuh",
    "gsr_forward_ws.cuh",
    "gsr_backward.cuh",
    "gsr_backward_region.cuh",
    "gsr_frontend.cuh",
    "gsr_loss.cuh",
    "gsr_umma.cuh",
    os.path.join("..", "..", "include", "gsraster.h"),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-cudart", "static",
    "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(defines, out_path: str) -> str:
    """Tuning aid: the same library with -D overrides of the geometry (gsr_common.cuh), written to
    out_path; select it at run time with GSR_LIB_PATH=out_path."""
    cmd = [nvcc_path()] + [f for f in NVCC_FLAGS if f != "-v" and f != "-Xptxas"] + [f"-D{d}" for d in defines]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out_path]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out_path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libgsraster.so for sm_100a; returns its path.  Safe when several processes call it at once (one rank
    per GPU under torchrun): the build runs under a file lock, into a per-process temporary, and whoever gets the lock
    second finds the library fresh."""
    if not force and not is_stale():
        return LIB
    import fcntl

    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB
            try:
                nvcc = nvcc_path()
            except RuntimeError:
                if os.path.exists(LIB):  # no compiler on this machine: the shipped library is what there is
                    sys.stderr.write("[gsasr_b200.build] sources are newer than libgsraster.so but nvcc is not here: using the library as is\n")
                    return LIB
                raise
            tmp = f"{LIB}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            log = proc.stdout + proc.stderr
            with open(os.path.join(HERE, "build.log"), "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if proc.returncode != 0:
                sys.stderr.write(log)
                raise RuntimeError("nvcc failed building libgsraster.so")
            os.replace(tmp, LIB)
            if verbose:
                sys.stderr.write(log)
            return LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
