"""Kernel LOGIC without a GPU.

The `-m gpu` parity tests are the proof of the CUDA path; they need a B200.  This test runs a subset of
those SAME test functions -- unchanged -- against the engine compiled for the CPU emulator of the CUDA
execution model (tests/emu/cuda_emu.h: every CUDA thread a fibre, barriers / shuffles / mbarriers / bulk
copies emulated), at reduced sizes (conftest.sz).  It catches indexing, prefix-sum and protocol bugs in
the kernels on the build machine; it says nothing about performance, and the emulated library is never
loaded by the product (ropebwt2_b200/binding.py only takes its path from tests/conftest.py under RB2_EMU=1)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SUBSET = ("test_resident_array_across_batches or test_regimes_alternate or test_forced_dense_multi_batch or test_edge_cases "
          "or (test_golden_fixtures and not batches) or test_rope_api or (test_uniform_one_batch and (1-2 or 0-3 or 2-4)) "
          "or (test_three_batches and 1-2) or (test_direct_delivery and 1) or (test_regimes_alternate_sharded and 2-2)")


def test_gpu_parity_tests_on_the_cpu_emulator():
    env = dict(os.environ, RB2_EMU="1", RB2_EMU_SMS="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_dense_regime_gpu.py"), os.path.join(ROOT, "tests", "test_parity_gpu.py"),
                        os.path.join(ROOT, "tests", "test_sharded_gpu.py"), "-m", "gpu", "-x", "-q", "-k", SUBSET, "-p", "no:cacheprovider"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=3000)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout, tail
