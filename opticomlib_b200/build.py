"""Build the sm_100a shared library IN-TREE with nvcc.

    python -m opticomlib_b200.build [--force]

Produces ``opticomlib_b200/_ssfm_b200.so`` (git-ignored, travels to the GPU box with the
snapshot).  nvcc cross-compiles without a GPU.  The static CUDA runtime is linked in, so the
library has no dependency beyond the driver and can be loaded from C, ctypes or next to PyTorch.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "_ssfm_b200.so")
SOURCES = ["ssfm_api.cu", "ssfm_wf_f64.cu", "ssfm_wf_f32.cu", "filtfilt.cu", "spectrum.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "ssfm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> str:
    """profile=True: a second library `_ssfm_b200_prof.so` with the clock64 phase instrumentation of k_wf
    (-DSSFM_WF_PROFILE; load it with SSFM_B200_LIB=...; experiments only, never the product path)."""
    if profile:
        return _build(os.path.join(HERE, "_ssfm_b200_prof.so"), os.path.join(HERE, "build", "prof"), ["-DSSFM_WF_PROFILE"], verbose)
    if not force and not _stale():
        return LIB
    return _build(LIB, os.path.join(HERE, "build"), [], verbose)


def _build(lib: str, objdir: str, extra: list, verbose: bool) -> str:
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, profile="--profile" in sys.argv))
