import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# RB2_EMU=1: run the tests marked `gpu` against the engine compiled for the CPU emulator of the CUDA
# execution model (tests/emu): same kernels, same host code, no GPU.  Sizes shrink through sz().
EMU = os.environ.get("RB2_EMU") == "1"


def sz(gpu: int, emu: int) -> int:
    """A test size: `gpu` on the real device, `emu` under the (much slower) CPU emulator."""
    return emu if EMU else gpu


not_on_emu = pytest.mark.skipif(EMU, reason="needs the real device / the real library")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if EMU:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        from ropebwt2_b200 import binding
        binding.load(path=build_emu.build())


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["cases"]


def flags_to_mode(flags: str):
    """reference CLI flags -> (sorting order, forward, reverse) as main.c:100-114 parses them."""
    so = 2 if "r" in flags else (1 if "s" in flags else 0)
    return so, "F" not in flags, "R" not in flags
