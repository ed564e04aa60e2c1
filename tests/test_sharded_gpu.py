"""GPU parity tests of the SHARDED build (one index over P ranks, include/ropebwt2_b200.h):
P "virtual ranks" -- P engines driven by P host threads on cuda:0, exchanging through the in-process
communicator -- must produce the BWT of the CPU oracle bit for bit, for every sorting order, with
N's, duplicates, variable lengths, empty rank shares and several batches (non-empty intervals).
The same SPMD code runs under NCCL with one process per GPU (tests/test_sharded_nccl.py)."""
import threading

import numpy as np
import pytest

from oracle import oracle as orc
from ropebwt2_b200 import load
from ropebwt2_b200.binding import ShardedEngine, local_group
from ropebwt2_b200.synth import encode_batch, genome_reads, uniform_reads, varlen_reads
from conftest import sz

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def split(reads, P):
    """contiguous shares, rank order = input order"""
    n = len(reads)
    cuts = [n * r // P for r in range(P + 1)]
    return [reads[cuts[r]:cuts[r + 1]] for r in range(P)]


class Cluster:
    def __init__(self, so, P, device=0):
        self.P = P
        self.grp = local_group(P)
        self.eng = [ShardedEngine(device, so, r, P, group=self.grp) for r in range(P)]

    def insert(self, bufs):
        errs = []

        def run(r):
            try:
                self.eng[r].insert_multi(bufs[r])
            except Exception as ex:  # pragma: no cover
                errs.append(ex)
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.P)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs

    def text(self):
        L = load()
        total = self.eng[0].total()
        blocks = []
        for s in range(36):
            o = L.rb2_shard_owner(self.P, s)
            for r in range(self.P):  # blocks of a sub-bucket exist on its owner only
                if r != o:
                    assert L.rb2_num_blocks(self.eng[r].h, s) == 0
            blocks.append(self.eng[o].fetch_subbucket(s))
        return orc.decode_blocks(np.concatenate(blocks), total)

    def close(self):
        for e in self.eng:
            e.close()
        load().rb2_group_destroy(self.grp)


def check(so, P, batches, fwd=True, rev=False):
    """batches: list of read collections; each is split over the ranks"""
    o = orc.Oracle(so)
    c = Cluster(so, P)
    for reads in batches:
        o.insert_multi(encode_batch(reads, fwd, rev))
        c.insert([encode_batch(part, fwd, rev) for part in split(reads, P)])
        for e in c.eng:  # every rank knows the whole-index marginals
            assert np.array_equal(e.counts(), o.counts())
    assert np.array_equal(c.text(), o.text()), f"so={so} P={P}"
    c.close()


@pytest.mark.parametrize("P", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("so", [0, 1, 2])
def test_uniform_one_batch(so, P):
    check(so, P, [uniform_reads(3000, 40, 7 + so, n_frac=0.01)])


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("so", [0, 1, 2])
def test_three_batches_nonempty_intervals(so, P):
    rd = genome_reads(6000, 50, 11, coverage=40.0)  # high coverage: many duplicate suffixes
    check(so, P, [rd[:2500], rd[2500:4000], rd[4000:]])


@pytest.mark.parametrize("P", [2, 8])
@pytest.mark.parametrize("so", [0, 1, 2])
def test_varlen_both_strands(so, P):
    check(so, P, [varlen_reads(700, 60, 5), varlen_reads(300, 20, 6)], fwd=True, rev=True)


def test_empty_shares_and_tiny_batches():
    for so in (0, 1):
        o = orc.Oracle(so)
        c = Cluster(so, 4)
        rd = uniform_reads(5, 12, 3)
        o.insert_multi(encode_batch(rd))
        c.insert([encode_batch(rd[:2]), np.zeros(0, np.uint8), encode_batch(rd[2:]), np.zeros(0, np.uint8)])
        one = uniform_reads(1, 30, 4)
        o.insert_multi(encode_batch(one))
        c.insert([np.zeros(0, np.uint8)] * 3 + [encode_batch(one)])
        assert np.array_equal(c.text(), o.text())
        c.close()


def test_long_runs_and_duplicates():
    """identical strings: group sizes stay large, counts above the 4-byte run limit never split wrongly"""
    rd = np.tile(uniform_reads(3, 25, 9), (4000, 1))
    check(1, 4, [rd])
    check(0, 2, [rd])


def test_medium_rlo_matches_single_gpu_engine():
    """200k x 101: the sharded index equals the one-GPU engine's (which is pinned to the reference)"""
    from ropebwt2_b200 import MRope
    rd = uniform_reads(200000, 101, 21)
    buf = encode_batch(rd)
    m = MRope(1)
    m.insert_multi(buf)
    want = orc.decode_index(load(), m.h, m.total())[0]
    m.close()
    c = Cluster(1, 4)
    c.insert([encode_batch(p) for p in split(rd, 4)])
    assert np.array_equal(c.text(), want)
    c.close()


@pytest.mark.parametrize("so", [0, 1, 2])
def test_direct_delivery_across_batches(so, monkeypatch):
    """Dense batches: the merge kernels store the new interval starts straight into the owner ranks' buffers
    (csrc/rb2_comm.h p2p_map).  The mappings are kept from batch to batch and re-made when a rank must grow
    (second batch larger), dropped for a send/recv batch (RB2_P2P=0) and made again afterwards."""
    monkeypatch.setenv("RB2_FLAT", "1")
    o = orc.Oracle(so)
    c = Cluster(so, 3)
    sizes = [sz(3000, 600), sz(9000, 1500), sz(2000, 400), sz(2500, 500), sz(2500, 500), sz(2500, 500), sz(12000, 2000)]
    for k, n in enumerate(sizes):
        rd = uniform_reads(n, 40, 50 + k, n_frac=0.01) if k != 2 else varlen_reads(n, 60, 52, 5)
        if k == 3:
            monkeypatch.setenv("RB2_P2P", "0")
        if k == 5:
            monkeypatch.setenv("RB2_FLAT", "0")  # a sparse batch in between: leaf blocks rebuilt, mappings dropped
        o.insert_multi(encode_batch(rd))
        c.insert([encode_batch(p) for p in split(rd, 3)])
        monkeypatch.delenv("RB2_P2P", raising=False)
        monkeypatch.setenv("RB2_FLAT", "1")
    st = c.eng[0].stats()
    assert st["flat_batches"] == len(sizes) - 1 and st["p2p_batches"] == len(sizes) - 2
    assert np.array_equal(c.text(), o.text())
    c.close()


@pytest.mark.parametrize("so", [0, 1])
def test_one_long_string_among_short_ones_sharded(so, monkeypatch):
    """A length outlier in a sharded batch: the ranks cut the batch together, in the global string
    order, instead of replicating a mostly empty symbol matrix on every rank."""
    from conftest import EMU
    if EMU:
        monkeypatch.setenv("RB2_SPLIT_SLACK", "4096")
    rng = np.random.default_rng(78)
    short = [rng.integers(1, 5, size=int(rng.integers(5, 40))).astype(np.uint8) for _ in range(sz(3000, 300))]
    contig = rng.integers(1, 5, size=sz(60_000, 2_000)).astype(np.uint8)
    strs = short[:len(short) // 3] + [contig] + short[len(short) // 3:]
    o = orc.Oracle(so)
    o.insert_multi(encode_batch(strs))
    c = Cluster(so, 3)
    c.insert([encode_batch(p) for p in split(strs, 3)])
    assert c.eng[0].stats()["n_columns"] > len(contig)  # (several sub-batches: more columns than the longest string)
    assert np.array_equal(c.text(), o.text())
    c.close()


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("so", [0, 1, 2])
def test_regimes_alternate_sharded(so, P, monkeypatch):
    """Sparse and dense batches alternating on a sharded index: the array is rebuilt from the leaf blocks of a non-empty
    index (sparse -> dense) and the blocks from the array (dense -> sparse) on every rank."""
    o = orc.Oracle(so)
    c = Cluster(so, P)
    for k, flat in enumerate(["0", "1", "1", "0", "1"]):
        monkeypatch.setenv("RB2_FLAT", flat)
        # (equal lengths: the last column of a sparse batch inserts every sentinel and splits leaf blocks, so the
        #  bucket table of the blocks changes behind the last control-block upload of that batch)
        rd = uniform_reads(sz(2500, 700) + 137 * k, sz(50, 30), 70 + k, n_frac=0.01)
        o.insert_multi(encode_batch(rd))
        c.insert([encode_batch(p) for p in split(rd, P)])
        for e in c.eng:
            assert np.array_equal(e.counts(), o.counts())
    assert c.eng[0].stats()["flat_batches"] == 3
    assert np.array_equal(c.text(), o.text())
    c.close()


@pytest.mark.parametrize("so", [0, 1])
def test_blocks_fetched_between_dense_batches_sharded(so, monkeypatch):
    """The leaf blocks of a sharded index are rebuilt on demand from the resident arrays (every fetch between two dense
    batches), the next batch goes on with the arrays, a reset starts over."""
    monkeypatch.setenv("RB2_FLAT", "1")
    o = orc.Oracle(so)
    c = Cluster(so, 3)
    for k in range(3):
        rd = uniform_reads(sz(3000, 500), sz(40, 25), 90 + k, n_frac=0.01)
        o.insert_multi(encode_batch(rd))
        c.insert([encode_batch(p) for p in split(rd, 3)])
        assert np.array_equal(c.text(), o.text()), k
    for e in c.eng:
        e.reset()
    o = orc.Oracle(so)
    rd = varlen_reads(sz(2000, 400), 50, 93, 1)
    o.insert_multi(encode_batch(rd))
    c.insert([encode_batch(p) for p in split(rd, 3)])
    assert np.array_equal(c.text(), o.text())
    c.close()
