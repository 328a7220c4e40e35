"""Build the C part of the oracle (test infrastructure) with gcc.

Outputs go to ``oracle/_build/`` (git-ignored, travels to the GPU box).
Called from ``__graft_entry__.build()``; building the checker is not using it.
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "filtfilt_oracle.c")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, src])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
