import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["cases"]


def flags_to_mode(flags: str):
    """reference CLI flags -> (sorting order, forward, reverse) as main.c:100-114 parses them."""
    so = 2 if "r" in flags else (1 if "s" in flags else 0)
    return so, "F" not in flags, "R" not in flags
