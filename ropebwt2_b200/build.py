"""Builds libropebwt2_b200.so in-tree (ropebwt2_b200/_build/): nvcc for the sm_100a engine,
gcc for the plain-C host API, one shared object that exports the C-ABI of include/*.h."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT = os.environ.get("RB2_BUILD_DIR", os.path.join(PKG, "_build"))
LIB = os.path.join(OUT, "libropebwt2_b200.so")
EXTRA = os.environ.get("RB2_NVCC_EXTRA", "").split()

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.write("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    if not force and _newer(LIB, srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cc = os.environ.get("CC", "gcc")
    with open(os.path.join(OUT, "build.log"), "w") as log:
        _run([nvcc] + NVCC_FLAGS + EXTRA + ["-c", os.path.join(CSRC, "rb2_engine.cu"), "-o", os.path.join(OUT, "rb2_engine.o")], log)
        _run([cc, "-O2", "-g", "-Wall", "-fPIC", "-c", os.path.join(CSRC, "mrope_b200.c"), "-o", os.path.join(OUT, "mrope_b200.o")], log)
        _run([nvcc, "-shared", "-o", LIB, os.path.join(OUT, "rb2_engine.o"), os.path.join(OUT, "mrope_b200.o"),
              "-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64"], log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
