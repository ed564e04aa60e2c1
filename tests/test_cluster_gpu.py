"""Several GPUs behind the unchanged mrope.h API (csrc/rb2_cluster.inl): with RB2_GPUS=P the engine
behind an mrope_t is a proxy over P sharded engines (ranks = threads of the process).  Here the P
ranks share cuda:0 (RB2_GPUS_SAME_DEVICE=1), so the test runs on a one-GPU box; on a node with P GPUs
the same code spreads over them.  Everything a user of the reference sees must be unchanged:
counts, the block iterator (= the BWT), mr_rank2a, mr_dump, and the reference's own driver."""
import os
import subprocess

import numpy as np
import pytest

from conftest import not_on_emu
from oracle import oracle as orc
from ropebwt2_b200 import MRope, load
from ropebwt2_b200.synth import encode_batch, genome_reads, reads_to_lines, varlen_reads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "ropebwt2_b200")


@pytest.fixture
def cluster4(monkeypatch):
    monkeypatch.setenv("RB2_GPUS", "4")
    monkeypatch.setenv("RB2_GPUS_SAME_DEVICE", "1")


def text(m):
    return orc.decode_index(load(), m.h, m.total())[0]


@pytest.mark.parametrize("so", [0, 1, 2])
def test_mrope_api_over_four_ranks(cluster4, so, tmp_path, monkeypatch):
    rd = genome_reads(8000, 60, 17 + so, coverage=40.0)
    o, m = orc.Oracle(so), MRope(so)
    for part in (rd[:5000], rd[5000:], varlen_reads(300, 40, 3)):
        buf = encode_batch(part, True, so == 2)
        o.insert_multi(buf)
        m.insert_multi(buf)
        assert np.array_equal(m.counts(), o.counts())
    want = o.text()
    assert np.array_equal(text(m), want)
    for x in (0, 1, 777, o.total() // 3, o.total() - 1, o.total()):
        assert np.array_equal(m.rank2a(x)[0], o.rank1a(x)), x
    # a dump written through the proxy is an ordinary .fmr: a one-GPU engine restores it
    path = str(tmp_path / "c.fmr")
    m.dump(path)
    m.close()
    monkeypatch.delenv("RB2_GPUS")
    one = MRope.restore(path)
    assert np.array_equal(text(one), want)
    one.close()


@not_on_emu  # (the drop-in binary links the real library)
@pytest.mark.skipif(not os.path.exists(DROPIN), reason="drop-in binary not built (oracle/Makefile: make dropin)")
def test_reference_driver_uses_all_ranks():
    """RB2_GPUS=4 ropebwt2_b200 -LRs: the reference's main.c, unmodified, on four ranks"""
    lines = reads_to_lines(genome_reads(9000, 101, 4))
    env = dict(os.environ, RB2_GPUS="4", RB2_GPUS_SAME_DEVICE="1")
    for flags in ("-LRs", "-LR", "-Lr", "-LRs -m 200k"):
        r = subprocess.run([DROPIN] + flags.split() + ["-"], input=lines, capture_output=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr.decode()[-500:]
        want, _ = orc.ref_cli(flags.split() + ["-"], lines)
        assert r.stdout == want, flags
