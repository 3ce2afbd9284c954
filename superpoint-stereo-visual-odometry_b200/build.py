"""Build libspvo_frontend.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libspvo_frontend.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """`make` in csrc/ (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...)."""
    cmd = ["make", "-C", CSRC, "-j4"] + (["-B"] if force else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libspvo_frontend.so failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
