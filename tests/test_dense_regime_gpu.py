"""GPU parity tests of the DENSE regime (csrc/rb2_flat.cuh): for the duration of a batch the BWT is a
flat symbol array rewritten by one streaming kernel per column, then re-encoded into leaf blocks.
The engine picks the regime per batch (RB2_FLAT=1 / 0 forces it); either way the decoded BWT must be
the oracle's bit for bit, and batches of the two regimes must chain (blocks -> flat -> blocks)."""
import numpy as np
import pytest

from conftest import not_on_emu, sz
from oracle import oracle as orc
from ropebwt2_b200 import MRope, load
from ropebwt2_b200.synth import encode_batch, genome_reads, uniform_reads, varlen_reads

pytestmark = pytest.mark.gpu


def text(m):
    return orc.decode_index(load(), m.h, m.total())[0]


@pytest.mark.parametrize("so", [0, 1, 2])
def test_forced_dense_multi_batch(monkeypatch, so):
    monkeypatch.setenv("RB2_FLAT", "1")
    n = sz(9000, 1800)
    rd = genome_reads(n, sz(60, 40), 13 + so, coverage=40.0)
    o, m = orc.Oracle(so), MRope(so)
    for part in (rd[:n * 4 // 9], rd[n * 4 // 9:n * 2 // 3], rd[n * 2 // 3:]):
        buf = encode_batch(part, True, so == 2)
        o.insert_multi(buf)
        m.insert_multi(buf)
        assert np.array_equal(text(m), o.text())
    assert m.stats()["flat_batches"] == 3
    assert np.array_equal(m.counts(), o.counts())
    m.close()


@pytest.mark.parametrize("so", [0, 1])
def test_regimes_alternate(monkeypatch, so):
    """dense, sparse, dense, sparse on one index: both conversions, non-empty intervals in both regimes"""
    o, m = orc.Oracle(so), MRope(so)
    for i, seed in enumerate((1, 2, 3, 4)):
        monkeypatch.setenv("RB2_FLAT", "1" if i % 2 == 0 else "0")
        buf = encode_batch(varlen_reads(sz(1500, 300), sz(70, 40), seed), True, True)
        o.insert_multi(buf)
        m.insert_multi(buf)
        assert np.array_equal(text(m), o.text()), (so, i)
    assert m.stats()["flat_batches"] == 2
    # the re-encoded pool is a legal index for the single-string path and for rank queries
    one = encode_batch(uniform_reads(1, 40, 9))
    o.insert_multi(one)
    m.insert1(one)
    assert np.array_equal(text(m), o.text())
    for x in (0, 1, o.total() // 3, o.total()):
        assert np.array_equal(m.rank2a(x)[0], o.rank1a(x))
    m.close()


@not_on_emu
def test_auto_choice_and_big_counts(monkeypatch):
    """without RB2_FLAT the engine picks dense for a short-read batch of this size; 300k copies of one
    read give per-symbol counts above the 4-byte run limit (records split, runs longer than a tile)"""
    monkeypatch.delenv("RB2_FLAT", raising=False)
    rd = np.concatenate([uniform_reads(150000, 101, 5), np.tile(uniform_reads(1, 101, 6), (600000, 1))])
    o, m = orc.Oracle(1), MRope(1)
    buf = encode_batch(rd)
    o.insert_multi(buf)
    m.insert_multi(buf)
    assert m.stats()["flat_batches"] == 1
    assert np.array_equal(text(m), o.text())
    m.close()


def test_batched_rank_queries(monkeypatch):
    """rb2_rank_batch (one warp per position) against the oracle's occ() on a re-encoded and on an
    in-place updated index"""
    rng = np.random.default_rng(5)
    for regime in ("1", "0"):
        monkeypatch.setenv("RB2_FLAT", regime)
        o, m = orc.Oracle(1), MRope(1)
        for seed in (3, 4):
            buf = encode_batch(genome_reads(sz(3000, 600), sz(70, 40), seed, coverage=30.0))
            o.insert_multi(buf)
            m.insert_multi(buf)
        xs = np.concatenate([[0, 1, o.total() - 1, o.total()], rng.integers(0, o.total() + 1, size=sz(3000, 400))])
        got = m.rank_batch(xs)
        for x, g in zip(xs[:300], got[:300]):
            assert np.array_equal(g, o.rank1a(int(x))), (regime, int(x))
        # consistency of the whole batch: occ is monotone and occ(total) are the marginals
        order = np.argsort(xs, kind="stable")
        assert (np.diff(got[order], axis=0) >= 0).all()
        assert np.array_equal(got[3], o.counts().sum(0))
        m.close()


@pytest.mark.parametrize("so", [0, 1, 2])
def test_resident_array_across_batches(monkeypatch, so):
    """Successive dense batches (the north-star shape: 12 mr_insert_multi calls into one growing index):
    the flat array stays resident, nothing is decoded in between, non-empty intervals from batch 2 on;
    leaf blocks are rebuilt only when the iterator asks for them at the end."""
    monkeypatch.setenv("RB2_FLAT", "1")
    n = sz(24000, 2400)
    rd = genome_reads(n, sz(60, 40), 23 + so, coverage=40.0)
    o, m = orc.Oracle(so), MRope(so)
    for k in range(6):
        buf = encode_batch(rd[k * n // 6:(k + 1) * n // 6], True, so == 2)
        o.insert_multi(buf)
        m.insert_multi(buf)
        assert np.array_equal(m.counts(), o.counts()), k   # mr_get_c is current after every call
    assert m.stats()["flat_batches"] == 6
    assert np.array_equal(text(m), o.text())
    for x in (0, 1, o.total() // 3, o.total()):
        assert np.array_equal(m.rank2a(x)[0], o.rank1a(x))
    # and the array is still usable afterwards
    buf = encode_batch(uniform_reads(sz(3000, 300), sz(60, 40), 5))
    o.insert_multi(buf)
    m.insert_multi(buf)
    assert np.array_equal(text(m), o.text())
    m.close()


def test_pipelined_calls_and_eager_counts(monkeypatch):
    """mr_insert_multi returns when the batch is on the device; the insertion runs behind it on a worker thread.
    The marginal counts (mr_get_c) must be right immediately -- they are computed from the batch itself -- and a
    reset queued between batches must take effect in order."""
    monkeypatch.setenv("RB2_FLAT", "1")
    n = sz(40000, 1500)
    rd = genome_reads(n, sz(80, 40), 31, coverage=30.0)
    o, m = orc.Oracle(1), MRope(1)
    for k in range(4):  # back to back, nothing in between waits for the worker
        buf = encode_batch(rd[k * n // 4:(k + 1) * n // 4])
        o.insert_multi(buf)
        m.insert_multi(buf)
        assert np.array_equal(m.counts(), o.counts()), k
    assert np.array_equal(text(m), o.text())
    # reset + new batches, queued behind each other
    L = load()
    o2 = orc.Oracle(1)
    L.rb2_reset(m.engine_handle)
    for k in (3, 1):
        buf = encode_batch(rd[k * n // 4:(k + 1) * n // 4])
        o2.insert_multi(buf)
        m.insert_multi(buf)
    assert np.array_equal(m.counts(), o2.counts())
    assert np.array_equal(text(m), o2.text())
    m.close()


@pytest.mark.parametrize("so", [0, 1])
def test_long_reads_in_the_dense_regime(monkeypatch, so):
    """Config-4 shape (many columns, all strings live until the end) forced through the dense regime, which is
    what the cost model picks for 1 M x 10 kbp: thousands of k_column_fused / k_flat_merge columns on one array."""
    monkeypatch.setenv("RB2_FLAT", "1")
    rd = uniform_reads(sz(300, 40), sz(4000, 300), 14)
    o, m = orc.Oracle(so), MRope(so)
    for part in (rd[:len(rd) * 2 // 3], rd[len(rd) * 2 // 3:]):
        buf = encode_batch(part)
        o.insert_multi(buf)
        m.insert_multi(buf)
    assert np.array_equal(m.counts(), o.counts())
    assert np.array_equal(text(m), o.text())
    m.close()


@pytest.mark.parametrize("so", [0, 1, 2])
def test_restore_then_dense_batches(monkeypatch, tmp_path, so):
    """BASELINE config 5 shape: mr_restore of a dumped index, then batches in the dense regime (the array is built from
    the restored leaf blocks), then a dump of the result that restores to the same text."""
    monkeypatch.setenv("RB2_FLAT", "1")
    n = sz(12000, 1500)
    rd = uniform_reads(n, sz(60, 30), 31 + so, n_frac=0.002)
    a, b = n // 2, n * 3 // 4
    o, m = orc.Oracle(so), MRope(so)
    buf = encode_batch(rd[:a], True, so == 2)
    o.insert_multi(buf)
    m.insert_multi(buf)
    p = str(tmp_path / "half.fmr")
    m.dump(p)
    m.close()
    m2 = MRope.restore(p)
    for part in (rd[a:b], rd[b:]):
        buf = encode_batch(part, True, so == 2)
        o.insert_multi(buf)
        m2.insert_multi(buf)
    assert m2.stats()["flat_batches"] == 2
    assert np.array_equal(m2.counts(), o.counts())
    assert np.array_equal(text(m2), o.text())
    q = str(tmp_path / "whole.fmr")
    m2.dump(q)
    m3 = MRope.restore(q)
    assert np.array_equal(text(m3), o.text())
    m2.close()
    m3.close()
