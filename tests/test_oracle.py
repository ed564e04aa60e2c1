"""CPU tests (no GPU): pin the oracle.

The C restatement (oracle/bcr_oracle.c) is checked against
  1. the golden fixtures produced by the unmodified reference binary (tests/golden/golden.json),
  2. the unmodified reference library oracle/_ref/libref.so on random multi-batch inputs,
  3. the README identities (README.md:18-25) through the reference binary,
  4. the naive suffix-sort definition (oracle/naive_bwt.py).
(2) and (3) need oracle/_ref, which is prebuilt in the build container and travels with the repo."""
import hashlib

import numpy as np
import pytest

from conftest import flags_to_mode
from oracle import oracle as orc
from oracle.naive_bwt import naive_bwt
from ropebwt2_b200.synth import (encode_batch, from_spec, reads_to_lines, text_to_ascii, uniform_reads)

needs_ref = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference once)")


def test_golden_cases_match_oracle(golden):
    for case in golden:
        so, fwd, rev = flags_to_mode(case["flags"])
        reads = from_spec(case["gen"])
        o = orc.Oracle(so)
        o.insert_multi(encode_batch(reads, fwd, rev))
        text = o.text()
        assert text.size == case["n_symbols"], case["name"]
        assert hashlib.md5(text_to_ascii(text)).hexdigest() == case["md5"], case["name"]
        assert np.bincount(text, minlength=6).tolist() == case["counts"], case["name"]
        if "text" in case:
            assert text_to_ascii(text)[:-1].decode() == case["text"]


def test_golden_batch_invariance(golden):
    """The text is independent of how the input is cut into batches."""
    for case in [c for c in golden if c["name"] in ("cfg1_rlo", "varlen_rclo_both", "cfg1N_io")]:
        so, fwd, rev = flags_to_mode(case["flags"])
        reads = from_spec(case["gen"])
        n = len(reads)
        o = orc.Oracle(so)
        for a, b in ((0, n // 3), (n // 3, n // 3 + 1), (n // 3 + 1, n)):
            o.insert_multi(encode_batch(reads[a:b], fwd, rev))
        assert hashlib.md5(text_to_ascii(o.text())).hexdigest() == case["md5"], case["name"]


def test_naive_definition():
    rng = np.random.default_rng(17)
    for it in range(60):
        so = it % 3
        strs = [rng.integers(1, 6 if it % 4 == 0 else 5, size=int(rng.integers(0, 8))).astype(np.uint8)
                for _ in range(int(rng.integers(1, 9)))]
        o = orc.Oracle(so)
        o.insert_multi(encode_batch(strs))
        assert np.array_equal(o.text(), naive_bwt(strs, so)), (it, [s.tolist() for s in strs])


@needs_ref
def test_oracle_vs_reference_library_random():
    rng = np.random.default_rng(23)
    for it in range(80):
        so = it % 3
        n = int(rng.integers(1, 40))
        strs = [rng.integers(1, 6 if it % 3 == 0 else 5, size=int(rng.integers(0, 30))).astype(np.uint8) for _ in range(n)]
        if it % 5 == 0:
            strs += [strs[0].copy(), strs[-1].copy()]
        rev = it % 2 == 1
        nb = 1 + it % 4
        cuts = sorted(rng.integers(0, len(strs) + 1, size=nb - 1).tolist())
        o, r = orc.Oracle(so), orc.RefLib(so)
        for a, b in zip([0] + cuts, cuts + [len(strs)]):
            if a == b:
                continue
            buf = encode_batch(strs[a:b], True, rev)
            o.insert_multi(buf)
            r.insert_multi(buf, 0)
        assert np.array_equal(o.text(), r.text()), it
        x = int(rng.integers(0, o.total() + 1))
        assert np.array_equal(o.rank1a(x), r.rank2a(x, -1)[0])


@needs_ref
def test_oracle_vs_reference_library_medium():
    for so in (0, 1, 2):
        rd = uniform_reads(4000, 60, 7 + so, n_frac=0.01)
        o, r = orc.Oracle(so), orc.RefLib(so)
        for a in (0, 1500, 3000):
            buf = encode_batch(rd[a:a + 1500], True, so != 1)
            o.insert_multi(buf)
            r.insert_multi(buf, 0)
        assert np.array_equal(o.text(), r.text())
        assert np.array_equal(o.counts().sum(0), np.bincount(r.text(), minlength=6))


@needs_ref
def test_readme_identities():
    """shuf | ropebwt2 -LRs == rev | sort | rev | ropebwt2 -LR, and the RCLO analogue (README.md:18-25)."""
    rd = uniform_reads(3000, 40, 5)
    lines = reads_to_lines(rd).split(b"\n")[:-1]
    rlo_sorted = sorted(lines, key=lambda s: s[::-1])
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rclo_sorted = sorted(lines, key=lambda s: s.translate(comp)[::-1])
    out_s, _ = orc.ref_cli(["-LRs", "-"], b"\n".join(lines) + b"\n")
    out_io, _ = orc.ref_cli(["-LR", "-"], b"\n".join(rlo_sorted) + b"\n")
    assert out_s == out_io
    out_r, _ = orc.ref_cli(["-LRr", "-"], b"\n".join(lines) + b"\n")
    out_io2, _ = orc.ref_cli(["-LR", "-"], b"\n".join(rclo_sorted) + b"\n")
    assert out_r == out_io2
    # and the oracle agrees with both sides
    o = orc.Oracle(1)
    o.insert_multi(encode_batch(rd))
    assert text_to_ascii(o.text()) == out_s
