"""TEST INFRASTRUCTURE: builds tests/emu/_build/libropebwt2_b200_emu.so -- the engine's kernels compiled
by g++ against the CPU emulator of the CUDA execution model (cuda_emu.h), for checking kernel LOGIC
on a machine without a GPU.  Never used by the product path or by any measurement."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ropebwt2_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libropebwt2_b200_emu.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))] + \
           [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    opt = os.environ.get("RB2_EMU_OPT", "-O2")
    common = ["-std=c++17"] + opt.split() + ["-g", "-fPIC", "-DRB2_EMU", "-I", os.path.join(HERE, "stubs"), "-include", os.path.join(HERE, "cuda_emu.h"), "-w"]
    cmds = [
        ["g++", "-x", "c++"] + common + ["-c", os.path.join(CSRC, "rb2_engine.cu"), "-o", os.path.join(OUT, "rb2_engine.o")],
        ["g++"] + common + ["-c", os.path.join(HERE, "cuda_emu.cpp"), "-o", os.path.join(OUT, "cuda_emu.o")],
        ["gcc", "-O2", "-g", "-fPIC", "-c", os.path.join(CSRC, "mrope_b200.c"), "-o", os.path.join(OUT, "mrope_b200.o")],
        ["g++", "-shared", "-o", LIB, os.path.join(OUT, "rb2_engine.o"), os.path.join(OUT, "cuda_emu.o"), os.path.join(OUT, "mrope_b200.o"), "-lpthread", "-ldl"],
    ]
    for c in cmds:
        r = subprocess.run(c, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("emulator build failed: " + " ".join(c))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
