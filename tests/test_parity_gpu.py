"""GPU parity tests: the CUDA engine, called through the reference-facing C-ABI (mrope.h / rope.h
as exported by libropebwt2_b200.so), against
  * the golden fixtures generated from the unmodified reference binary,
  * the CPU oracle (oracle/bcr_oracle.c) and the reference library (oracle/_ref/libref.so) on the
    same seeded inputs,
  * size-independent properties at larger sizes (batch invariance, RLO order invariance, symbol
    conservation).
Bit-exact: every comparison is on the decoded BWT symbol sequence (integer/byte work, no tolerance)."""
import ctypes as C
import hashlib
import os
import tempfile

import numpy as np
import pytest

from conftest import EMU, flags_to_mode, not_on_emu, sz
from oracle import oracle as orc
from ropebwt2_b200 import MRope, load
from ropebwt2_b200.synth import (encode_batch, from_spec, genome_reads, text_to_ascii, uniform_reads, varlen_reads)

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")


def gpu_text(m: MRope, total: int = None) -> np.ndarray:
    total = m.total() if total is None else total
    return orc.decode_index(load(), m.h, total)[0]


def md5_ascii(text: np.ndarray) -> str:
    return hashlib.md5(text_to_ascii(text)).hexdigest()


def build_both(so, batches):
    o, m = orc.Oracle(so), MRope(so)
    for buf in batches:
        o.insert_multi(buf)
        m.insert_multi(buf)
    return o, m


def test_golden_fixtures(golden):
    for case in golden:
        if EMU and case["n_symbols"] > 200000:
            continue
        so, fwd, rev = flags_to_mode(case["flags"])
        m = MRope(so)
        m.insert_multi(encode_batch(from_spec(case["gen"]), fwd, rev))
        assert m.total() == case["n_symbols"], case["name"]
        text = gpu_text(m)
        assert md5_ascii(text) == case["md5"], case["name"]
        assert m.counts().sum(0).tolist() == case["counts"], case["name"]
        m.close()


def test_golden_fixtures_in_batches(golden):
    """Same md5 when the input arrives in several mr_insert_multi calls (main.c:238-251 flushes)."""
    for case in golden:
        if EMU and case["n_symbols"] > 200000:
            continue
        so, fwd, rev = flags_to_mode(case["flags"])
        reads = from_spec(case["gen"])
        n = len(reads)
        m = MRope(so)
        cuts = [0, n // 5, n // 5 + 1, n // 2, n]
        for a, b in zip(cuts[:-1], cuts[1:]):
            m.insert_multi(encode_batch(reads[a:b], fwd, rev))
        assert md5_ascii(gpu_text(m)) == case["md5"], case["name"]
        m.close()


def test_random_small_vs_oracle():
    rng = np.random.default_rng(101)
    for it in range(sz(120, 24)):
        so = it % 3
        n = int(rng.integers(1, 30))
        strs = [rng.integers(1, 6 if it % 4 == 0 else 5, size=int(rng.integers(0, 25))).astype(np.uint8) for _ in range(n)]
        if it % 5 == 0:
            strs += [strs[0].copy(), strs[-1].copy(), strs[0].copy()]
        nb = 1 + it % 4
        cuts = sorted(rng.integers(0, len(strs) + 1, size=nb - 1).tolist())
        bufs = [encode_batch(strs[a:b], True, it % 2 == 1) for a, b in zip([0] + cuts, cuts + [len(strs)]) if b > a]
        o, m = build_both(so, bufs)
        assert np.array_equal(m.counts(), o.counts()), it
        assert np.array_equal(gpu_text(m), o.text()), (it, so, [s.tolist() for s in strs], cuts)
        m.close()


@pytest.mark.parametrize("so", [0, 1, 2])
def test_edge_cases(so):
    z = np.zeros(1, dtype=np.uint8)
    cases = {
        "one empty string": [z],
        "only empty strings": [np.zeros(7, dtype=np.uint8)],
        "one symbol": [np.array([3, 0], dtype=np.uint8)],
        "all N": [encode_batch([np.full(9, 5, dtype=np.uint8)] * 4)],
        "identical strings": [encode_batch([np.array([1, 2, 3, 4, 1, 1], dtype=np.uint8)] * 50, True, True)],
        "empty then non-empty batches": [np.zeros(3, dtype=np.uint8), encode_batch(varlen_reads(40, 12, 3)), np.zeros(2, dtype=np.uint8)],
        "ragged": [encode_batch(varlen_reads(500, 70, 8), True, False), encode_batch(varlen_reads(300, 5, 9), True, True)],
    }
    for name, bufs in cases.items():
        o, m = build_both(so, bufs)
        assert np.array_equal(m.counts(), o.counts()), name
        assert np.array_equal(gpu_text(m), o.text()), name
        m.close()


@pytest.mark.parametrize("so", [1, 2])
def test_runs_longer_than_the_4_byte_form(so):
    """600k copies of one read give runs > 2^19 symbols, which the device stores as several
    adjacent runs; the decoded text must not change."""
    rd = np.tile(np.array([[1, 2, 2, 4]], dtype=np.uint8), (sz(600000, 530000), 1))
    buf = encode_batch(rd)
    o, m = build_both(so, [buf, encode_batch(uniform_reads(100, 6, 1))])
    assert np.array_equal(gpu_text(m), o.text())
    m.close()


@pytest.mark.parametrize("so", [0, 1, 2])
def test_medium_multi_batch_vs_oracle(so):
    n, ln = sz(30000, 1500), sz(80, 40)
    rd = uniform_reads(n, ln, 40 + so, n_frac=0.002)
    bufs = [encode_batch(rd[a:a + n // 3], True, so == 2) for a in range(0, n, n // 3)]
    o, m = build_both(so, bufs)
    assert np.array_equal(m.counts(), o.counts())
    assert np.array_equal(gpu_text(m), o.text())
    rng = np.random.default_rng(so)
    for x in rng.integers(0, o.total() + 1, size=20).tolist() + [0, o.total()]:
        assert np.array_equal(m.rank2a(int(x))[0], o.rank1a(int(x))), x
    x, y = sorted(rng.integers(0, o.total() + 1, size=2).tolist())
    cx, cy = m.rank2a(x, y)
    assert np.array_equal(cx, o.rank1a(x)) and np.array_equal(cy, o.rank1a(y))
    m.close()


@pytest.mark.parametrize("so", [0, 1])
def test_long_reads_vs_oracle(so):
    """Long-string path (BASELINE config 4 shape, scaled): many columns, few strings per column."""
    rd = uniform_reads(60, sz(6000, 300), 4)
    o, m = build_both(so, [encode_batch(rd[:40]), encode_batch(rd[40:])])
    assert np.array_equal(gpu_text(m), o.text())
    m.close()


@needs_ref
@pytest.mark.parametrize("so", [0, 1, 2])
def test_vs_reference_library(so):
    """Directly against the unmodified reference mr_insert_multi (libref.so), realistic reads."""
    n = sz(60000, 2400)
    rd = genome_reads(n, sz(101, 50), 3 + so)
    r, m = orc.RefLib(so), MRope(so)
    for a in range(0, n, n * 5 // 12):
        buf = encode_batch(rd[a:a + n * 5 // 12], True, so == 2)
        r.insert_multi(buf, 1)
        m.insert_multi(buf)
    assert np.array_equal(gpu_text(m), r.text())
    for x in (0, 1, min(12345, r.total()), r.total() // 2, r.total()):
        assert np.array_equal(m.rank2a(x)[0], r.rank2a(x, -1)[0])
    m.close()


def test_rlo_is_input_order_invariant_at_scale():
    """README.md:18-25: the RLO BWT does not depend on the input order.  2M x 101 bp, plus symbol
    conservation (every base and one sentinel per string ends up in the BWT)."""
    rd = uniform_reads(sz(2_000_000, 3000), sz(101, 40), 2)
    m1 = MRope(1)
    m1.insert_multi(encode_batch(rd))
    perm = np.random.default_rng(0).permutation(rd.shape[0])
    m2 = MRope(1)
    half = rd.shape[0] // 2
    m2.insert_multi(encode_batch(rd[perm[:half]]))
    m2.insert_multi(encode_batch(rd[perm[half:]]))
    c1 = m1.counts()
    assert np.array_equal(c1, m2.counts())
    hist = np.bincount(rd.reshape(-1), minlength=6)
    hist[0] = rd.shape[0]
    assert np.array_equal(c1.sum(0), hist)
    # bucket b holds the symbols followed by b: its size is the number of b's (one $ per string)
    assert np.array_equal(c1.sum(1), hist)
    t1, t2 = gpu_text(m1), gpu_text(m2)
    assert hashlib.md5(t1.tobytes()).hexdigest() == hashlib.md5(t2.tobytes()).hexdigest()
    m1.close()
    m2.close()


def test_insert1_matches_insert_multi():
    strs = varlen_reads(25, 14, 5)
    for so in (0, 1, 2):
        m1, m2 = MRope(so), MRope(so)
        for s in strs:
            m1.insert1(np.concatenate([s[::-1], np.zeros(1, dtype=np.uint8)]))
        m2.insert_multi(encode_batch(strs))
        assert np.array_equal(gpu_text(m1), gpu_text(m2)), so
        if orc.have_ref():
            r = orc.RefLib(so)
            m3 = MRope(so)
            for s in strs:
                b = np.concatenate([s[::-1], np.zeros(1, dtype=np.uint8)])
                want = r.lib().mr_insert1(r.h, b.ctypes.data_as(C.POINTER(C.c_uint8)))
                assert m3.insert1(b) == want
        m1.close()
        m2.close()


def test_fmr_dump_restore_roundtrip(tmp_path):
    n = sz(20000, 1500)
    cut = n * 3 // 5
    rd = uniform_reads(n, sz(60, 40), 12, n_frac=0.001)
    for so in (0, 1):
        m = MRope(so)
        m.insert_multi(encode_batch(rd[:cut]))
        p = str(tmp_path / f"half{so}.fmr")
        m.dump(p)
        m2 = MRope.restore(p)
        assert m2.struct.so == so
        assert np.array_equal(m2.counts(), m.counts())
        assert np.array_equal(gpu_text(m2), gpu_text(m))
        # keep inserting into the restored index == one shot (BASELINE config 5 shape)
        m2.insert_multi(encode_batch(rd[cut:]))
        one = MRope(so)
        one.insert_multi(encode_batch(rd))
        assert np.array_equal(gpu_text(m2), gpu_text(one))
        for x in (m, m2, one):
            x.close()


@needs_ref
def test_fmr_interop_with_reference_binary(tmp_path):
    """Our .fmr is readable by the reference's -i (and it can keep inserting into it); the
    reference's -b dump is readable by mr_restore."""
    from ropebwt2_b200.synth import reads_to_lines
    n = sz(8000, 1200)
    h = n * 5 // 8
    rd = uniform_reads(n, sz(50, 30), 21)
    for so, flag in ((0, ""), (1, "s")):
        m = MRope(so)
        m.insert_multi(encode_batch(rd[:h]))
        ours = str(tmp_path / f"ours{so}.fmr")
        m.dump(ours)
        # reference re-emits our dump unchanged (FASTA mode on an empty input: SURVEY.md section 4 quirk)
        out, _ = orc.ref_cli(["-i", ours, "/dev/null"])
        assert out == text_to_ascii(gpu_text(m))
        # reference continues inserting into our dump == reference one-shot
        out2, _ = orc.ref_cli(["-LR", "-i", ours, "-"], reads_to_lines(rd[h:]))
        ref_one, _ = orc.ref_cli(["-LR" + flag, "-"], reads_to_lines(rd))
        assert out2 == ref_one
        # reference dump -> our restore -> continue inserting == reference one-shot
        theirs = str(tmp_path / f"ref{so}.fmr")
        dump, _ = orc.ref_cli(["-LRb" + flag, "-"], reads_to_lines(rd[:h]))
        open(theirs, "wb").write(dump)
        m3 = MRope.restore(theirs)
        m3.insert_multi(encode_batch(rd[h:]))
        assert text_to_ascii(gpu_text(m3)) == ref_one
        m.close()
        m3.close()


def test_rope_api():
    """rope.h: rope_insert_run returns rank(a, x) before the insertion; rope_rank2a counts."""
    L = load()
    rope = L.rope_init(64, 512)
    rng = np.random.default_rng(3)
    model = []
    for it in range(60):
        x = int(rng.integers(0, len(model) + 1))
        a = int(rng.integers(0, 6))
        rl = int(rng.integers(1, 40)) if it % 10 else int(rng.integers(1000, 3000))
        want = model[:x].count(a)
        got = L.rope_insert_run(rope, x, a, rl, None)
        assert got == want, (it, x, a, rl)
        model[x:x] = [a] * rl
    cx = np.zeros(6, dtype=np.int64)
    cy = np.zeros(6, dtype=np.int64)
    i64p = C.POINTER(C.c_int64)
    for _ in range(10):
        x, y = sorted(rng.integers(0, len(model) + 1, size=2).tolist())
        L.rope_rank2a(rope, x, y, cx.ctypes.data_as(i64p), cy.ctypes.data_as(i64p))
        assert cx.tolist() == [model[:x].count(a) for a in range(6)]
        assert cy.tolist() == [model[:y].count(a) for a in range(6)]
    L.rope_destroy(rope)


DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ropebwt2_b200")


@needs_ref
@not_on_emu
@pytest.mark.skipif(not os.path.exists(DROPIN), reason="drop-in binary not built (oracle/Makefile: make dropin)")
def test_reference_driver_on_our_library(tmp_path):
    """The reference's own main.c / rld0.c / crlf.c linked against libropebwt2_b200.so instead of
    its mrope.c / rope.c / rle.c: text, .fmd (-d) and CRLF (-B) outputs are byte-identical to the
    reference binary's; .fmr (-b) round-trips through the reference."""
    import subprocess
    from ropebwt2_b200.synth import reads_to_lines
    rd = genome_reads(12000, 101, 9)
    lines = reads_to_lines(rd)

    def ours(args):
        r = subprocess.run([DROPIN] + args, input=lines, capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr.decode()[-500:]
        return r.stdout

    for flags in ("-LR", "-LRs", "-Lr", "-LRsd", "-LRB", "-Ld", "-LRs -m 300k"):
        want, _ = orc.ref_cli(flags.split() + ["-"], lines)
        assert ours(flags.split() + ["-"]) == want, flags
    fmr = str(tmp_path / "x.fmr")
    open(fmr, "wb").write(ours(["-LRbs", "-"]))
    out, _ = orc.ref_cli(["-i", fmr, "/dev/null"])
    want, _ = orc.ref_cli(["-LRs", "-"], lines)
    assert out == want


@pytest.mark.parametrize("so", [0, 1])
def test_one_long_string_among_short_ones(so, monkeypatch):
    """A length outlier must not blow up the per-batch state: the engine cuts such a batch into
    string ranges; the reference accepts any mix of lengths, and the BWT does not depend on the cut."""
    if EMU:
        monkeypatch.setenv("RB2_SPLIT_SLACK", "4096")  # (the reduced sizes of the emulator still take the cutting path)
    rng = np.random.default_rng(77)
    short = [rng.integers(1, 5, size=int(rng.integers(5, 40))).astype(np.uint8) for _ in range(sz(3000, 400))]
    contig = rng.integers(1, 5, size=sz(150_000, 2_500)).astype(np.uint8)
    strs = short[:len(short) // 3] + [contig] + short[len(short) // 3:]
    o, m = build_both(so, [encode_batch(strs)])
    assert np.array_equal(m.counts(), o.counts())
    assert np.array_equal(gpu_text(m), o.text())
    m.close()


@not_on_emu
@pytest.mark.parametrize("flags", ["-LRs", "-LR", "-LRr"])
def test_beyond_2_32_symbols_md5_vs_reference(flags):
    """45 M x 101 bp = 4.59 G symbols (> 2^32 positions) in one batch, RLO / input order / RCLO: md5 of the decoded
    index against the md5 of the UNMODIFIED reference's output on the same seeded reads (tests/golden/
    ref_full_runs.json, recorded by tools/ref_full_run.py --workload cfg2 --reads 45000000 --flags ...).  Reads are
    generated on the GPU with the counter-based generator that tests/test_synth.py pins to the one that fed the reference."""
    import torch
    from ropebwt2_b200 import synth
    w = synth.workload("cfg2", 45_000_000)
    rec = orc.ref_recorded(w, flags)
    if rec is None:
        pytest.skip("no recorded reference run for " + synth.workload_key(w, flags))
    so, fwd, rev = flags_to_mode(flags)
    assert fwd and not rev
    nbytes = w["n"] * (w["L"] + 1)
    host = np.empty(nbytes, dtype=np.uint8)
    ht = torch.from_numpy(host)
    step = 3_000_000
    for a in range(0, w["n"], step):
        b = min(w["n"], a + step)
        t = torch.empty((b - a) * (w["L"] + 1), dtype=torch.uint8, device="cuda")
        synth.fill_batch_torch(t, w, a, b)
        ht[a * (w["L"] + 1):b * (w["L"] + 1)].copy_(t)
        del t
    torch.cuda.empty_cache()
    m = MRope(so)
    m.insert_multi(host)
    assert m.total() == nbytes
    md5, total = orc.index_md5(load(), m.h)
    assert total == nbytes and md5 == rec["md5_text"]
    m.close()
