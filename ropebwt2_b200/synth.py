"""Seeded synthetic read sets and the host-side batch encoding of the reference driver.

This is the host mirror of the *caller* side of the hot path: reference ``main.c:177-237``
turns every read into nt6 codes ($=0 A=1 C=2 G=3 T=4 N=5, main.c:17-26), reverses it,
and appends "reversed string + NUL" (and optionally the reversed reverse-complement,
main.c:227-236) to one buffer that is handed to ``mr_insert_multi`` (main.c:238-251).
``encode_batch`` produces exactly that buffer with numpy so tests and ``bench.py`` can
feed the C-ABI the same bytes the reference driver would.

Generators follow SURVEY.md section 8(d): **U** = iid uniform ACGT, **G** = reads sampled
from a random genome at a given coverage with substitution errors, both from
``numpy.random.default_rng(seed)``.
"""
from __future__ import annotations

import numpy as np

NT6 = np.frombuffer(b"$ACGTN", dtype=np.uint8)


def uniform_reads(n: int, length: int, seed: int, n_frac: float = 0.0) -> np.ndarray:
    """(n, length) uint8 matrix of nt6 codes 1..4 (plus N=5 with probability ``n_frac``)."""
    rng = np.random.default_rng(seed)
    r = rng.integers(1, 5, size=(n, length), dtype=np.uint8)
    if n_frac > 0:
        r[rng.random((n, length)) < n_frac] = 5
    return r


def genome_reads(n: int, length: int, seed: int, coverage: float = 30.0,
                 err: float = 0.01) -> np.ndarray:
    """Reads at uniform positions of a random genome (size n*length/coverage), random
    strand, ``err`` substitution rate.  Returns (n, length) uint8 nt6 codes 1..4."""
    rng = np.random.default_rng(seed)
    glen = max(length + 1, int(n * length / coverage))
    genome = rng.integers(1, 5, size=glen, dtype=np.uint8)
    start = rng.integers(0, glen - length + 1, size=n)
    idx = start[:, None] + np.arange(length)[None, :]
    r = genome[idx]
    strand = rng.random(n) < 0.5
    r[strand] = (5 - r[strand])[:, ::-1]
    sub = rng.random((n, length)) < err
    # substitute with one of the three other bases
    r[sub] = ((r[sub] - 1 + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) % 4) + 1
    return np.ascontiguousarray(r)


def varlen_reads(n: int, max_len: int, seed: int, min_len: int = 0):
    """Variable-length reads (list of 1-D arrays) with some N's and two duplicated strings."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ln = int(rng.integers(min_len, max_len + 1))
        out.append(rng.integers(1, 6 if rng.random() < 0.2 else 5, size=ln).astype(np.uint8))
    out += [out[0].copy(), out[len(out) // 2].copy()]
    return out


def from_spec(gen: dict):
    """Reads from a fixture spec: {"kind": "U"|"G"|"V", "n", "L", "seed", ...} (tests/golden)."""
    k = gen["kind"]
    if k == "U":
        return uniform_reads(gen["n"], gen["L"], gen["seed"], gen.get("n_frac", 0.0))
    if k == "G":
        return genome_reads(gen["n"], gen["L"], gen["seed"], gen.get("coverage", 30.0), gen.get("err", 0.01))
    if k == "V":
        return varlen_reads(gen["n"], gen["L"], gen["seed"], gen.get("lmin", 0))
    raise ValueError(k)


def revcomp(reads: np.ndarray) -> np.ndarray:
    """Reverse complement of fixed-length nt6 reads (A<->T, C<->G; $ and N unchanged)."""
    r = reads[:, ::-1].copy()
    m = (r >= 1) & (r <= 4)
    r[m] = 5 - r[m]
    return r


def encode_batch(reads, forward: bool = True, reverse: bool = False) -> np.ndarray:
    """The byte buffer reference main.c builds for ``mr_insert_multi``.

    ``reads`` is an (n, L) uint8 matrix or a list of 1-D uint8 arrays (variable length)
    in nt6 codes.  For every read, in order: the reversed read + NUL if ``forward``
    (main.c:200-203, 223-225), then the reversed reverse-complement + NUL if ``reverse``
    (main.c:227-236).  Returns a 1-D uint8 array ending in NUL (mrope.c:268)."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, length = reads.shape
        parts = []
        if forward:
            parts.append(reads[:, ::-1])
        if reverse:
            parts.append(revcomp(reads)[:, ::-1])
        k = len(parts)
        out = np.zeros((n, k, length + 1), dtype=np.uint8)
        for i, p in enumerate(parts):
            out[:, i, :length] = p
        return out.reshape(-1)
    chunks = []
    z = np.zeros(1, dtype=np.uint8)
    for r in reads:
        r = np.asarray(r, dtype=np.uint8)
        if forward:
            chunks += [r[::-1], z]
        if reverse:
            rc = r[::-1].copy()
            m = (rc >= 1) & (rc <= 4)
            rc[m] = 5 - rc[m]
            chunks += [rc[::-1], z]
    if not chunks:
        return np.zeros(0, dtype=np.uint8)
    return np.concatenate(chunks)


def reads_to_lines(reads) -> bytes:
    """One-sequence-per-line text (the ``-L`` input format, main.c:180-186)."""
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        n, length = reads.shape
        buf = np.empty((n, length + 1), dtype=np.uint8)
        buf[:, :length] = NT6[reads]
        buf[:, length] = ord("\n")
        return buf.tobytes()
    return b"".join(NT6[np.asarray(r, dtype=np.uint8)].tobytes() + b"\n" for r in reads)


def text_to_ascii(text: np.ndarray) -> bytes:
    """nt6 codes -> the characters the reference prints (main.c:308-313) + newline."""
    return NT6[text].tobytes() + b"\n"
