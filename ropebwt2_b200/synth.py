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


# ---------------------------------------------------------------------------------------------
# Counter-based generators for the full-size workloads (BASELINE.json configs 2-5).
#
# Every base is a pure function of (seed, read index, offset), built from the splitmix64 finaliser
# in wrapping 64-bit integer arithmetic.  That makes a read set (a) reproducible chunk by chunk in
# any order -- the 122 GB of config 3 never exists in one piece -- and (b) bit-identical between
# numpy on a CPU (tools/ref_full_run.py feeds the reference binary here) and torch on the GPU
# (bench.py fills the batches on the B200 in seconds); tests/test_synth.py pins the two together.
# ---------------------------------------------------------------------------------------------
_M1, _M2, _GAMMA = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0x9E3779B97F4A7C15
_K_GENOME, _K_START, _K_STRAND, _K_ERR, _K_BASE = 0x1000003D1, 0x2000005A7, 0x30000071B, 0x400000963, 0x500000B3F

WORKLOADS = {
    # SURVEY.md section 8(d); flags are the reference command line the md5 is taken with
    "cfg2": {"kind": "U", "n": 100_000_000, "L": 101, "seed": 2},
    "cfg3": {"kind": "G", "n": 1_200_000_000, "L": 101, "seed": 3, "coverage": 30.0, "err": 0.01},
    "cfg3u": {"kind": "U", "n": 1_200_000_000, "L": 101, "seed": 3},
    "cfg4": {"kind": "U", "n": 1_000_000, "L": 10_000, "seed": 4},
    "cfg5base": {"kind": "G", "n": 1_000_000_000, "L": 101, "seed": 3, "coverage": 30.0, "err": 0.01},
    "cfg5add": {"kind": "G", "n": 200_000_000, "L": 101, "seed": 5, "coverage": 30.0, "err": 0.01,
                "genome_seed": 3, "genome_n": 1_000_000_000},
}


def workload(name: str, reads: int = 0) -> dict:
    """A workload dict from WORKLOADS (optionally with another number of reads) or from
    "kind:n:L:seed".  For kind G the genome length follows the number of reads (30x coverage)."""
    if name in WORKLOADS:
        w = dict(WORKLOADS[name])
    else:
        k, n, L, seed = name.split(":")
        w = {"kind": k, "n": int(n), "L": int(L), "seed": int(seed)}
        if k == "G":
            w.update(coverage=30.0, err=0.01)
    if reads:
        if "genome_n" in w:
            w["genome_n"] = max(1, int(w["genome_n"] * reads / w["n"]))
        w["n"] = reads
    return w


def workload_key(w: dict, flags: str) -> str:
    return "%s n=%d L=%d seed=%d%s %s" % (w["kind"], w["n"], w["L"], w["seed"],
                                           (" cov=%g err=%g" % (w["coverage"], w["err"]) if w["kind"] == "G" else "")
                                           + (" genome=%d/%d" % (w["genome_seed"], w["genome_n"]) if "genome_n" in w else ""), flags)


def _signed(c: int) -> int:
    return c - (1 << 64) if c >= (1 << 63) else c


def _mix_np(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    z = x + np.uint64(_GAMMA)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
    return z ^ (z >> np.uint64(31))


def _mix_torch(x):
    """The same on an int64 torch tensor: wrapping multiply, logical shifts by masking."""
    z = x + _signed(_GAMMA)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _signed(_M1)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _signed(_M2)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


def _glen(w: dict) -> int:
    n = w.get("genome_n", w["n"])
    return max(w["L"] + 1, int(n * w["L"] / w["coverage"]))


def hash_reads(w: dict, a: int, b: int, device=None):
    """Reads a..b-1 of workload `w` as a (b-a, L) uint8 matrix of nt6 codes 1..4, forward
    orientation.  device=None: numpy array; otherwise a torch tensor on that device."""
    L, seed = w["L"], w["seed"]
    if device is None:
        with np.errstate(over="ignore"):
            u = np.uint64
            r = np.arange(a, b, dtype=np.uint64)[:, None]
            j = np.arange(L, dtype=np.uint64)[None, :]
            if w["kind"] == "U":
                return (u(1) + (_mix_np(u(seed * _K_BASE) + r * u(L) + j) >> u(62))).astype(np.uint8)
            glen, gs = _glen(w), w.get("genome_seed", seed)
            start = (_mix_np(u(seed * _K_START) + r) >> u(1)) % u(glen - L + 1)
            rev = (_mix_np(u(seed * _K_STRAND) + r) >> u(63)).astype(bool)
            pos = np.where(rev, start + u(L - 1) - j, start + j)
            base = (u(1) + (_mix_np(u(gs * _K_GENOME) + pos) >> u(62))).astype(np.uint8)
            base = np.where(rev, np.uint8(5) - base, base)
            e = _mix_np(u(seed * _K_ERR) + r * u(L) + j)
            sub = (e & u(0xFFFFFF)) < u(int(w["err"] * (1 << 24)))
            shift = (u(1) + ((e >> u(24)) & u(0xFFFF)) % u(3)).astype(np.uint8)
            return np.where(sub, (base - np.uint8(1) + shift) % np.uint8(4) + np.uint8(1), base).astype(np.uint8)
    import torch
    r = torch.arange(a, b, dtype=torch.int64, device=device)[:, None]
    j = torch.arange(L, dtype=torch.int64, device=device)[None, :]

    def top(z, bits):
        return (z >> (64 - bits)) & ((1 << bits) - 1)
    if w["kind"] == "U":
        return (1 + top(_mix_torch(_signed((seed * _K_BASE) & ((1 << 64) - 1)) + r * L + j), 2)).to(torch.uint8)
    glen, gs = _glen(w), w.get("genome_seed", seed)
    start = ((_mix_torch(seed * _K_START + r) >> 1) & ((1 << 63) - 1)) % (glen - L + 1)
    rev = top(_mix_torch(seed * _K_STRAND + r), 1).bool()
    pos = torch.where(rev, start + (L - 1) - j, start + j)
    base = 1 + top(_mix_torch(gs * _K_GENOME + pos), 2)
    base = torch.where(rev, 5 - base, base)
    e = _mix_torch(seed * _K_ERR + r * L + j)
    sub = (e & 0xFFFFFF) < int(w["err"] * (1 << 24))
    shift = 1 + ((e >> 24) & 0xFFFF) % 3
    return torch.where(sub, (base - 1 + shift) % 4 + 1, base).to(torch.uint8)


def stream_lines(w: dict, chunk: int = 2_000_000, threads: int = 2):
    """The workload as `-L` text (one read per line, main.c:180-186), chunk by chunk (bytes)."""
    lib = c_generator()
    chunk = max(1, min(chunk, 256_000_000 // (w["L"] + 1)))  # (bounded in bytes: long reads)
    for a in range(0, w["n"], chunk):
        b = min(w["n"], a + chunk)
        if lib is not None:
            yield c_reads(lib, w, a, b, True, threads).tobytes()
            continue
        r = hash_reads(w, a, b)
        buf = np.empty((b - a, w["L"] + 1), dtype=np.uint8)
        buf[:, :w["L"]] = NT6[r]
        buf[:, w["L"]] = 10
        yield buf.tobytes()


def fill_batch_np(dst: np.ndarray, w: dict, a: int, b: int, chunk: int = 2_000_000) -> None:
    """The mr_insert_multi buffer of reads a..b-1 (reversed read + NUL each, main.c:200-225)."""
    view = dst.reshape(b - a, w["L"] + 1)
    for x in range(a, b, chunk):
        y = min(b, x + chunk)
        view[x - a:y - a, :w["L"]] = hash_reads(w, x, y)[:, ::-1]
        view[x - a:y - a, w["L"]] = 0


def fill_batch_torch(dst, w: dict, a: int, b: int, chunk: int = 8_000_000) -> None:
    """The same into a 1-D uint8 torch tensor on a GPU (dst.numel() == (b-a)*(L+1))."""
    import torch
    chunk = max(1, min(chunk, 400_000_000 // (w["L"] + 1)))  # the generator works on int64 temporaries: bound them
    view = dst.view(b - a, w["L"] + 1)
    for x in range(a, b, chunk):
        y = min(b, x + chunk)
        view[x - a:y - a, :w["L"]] = torch.flip(hash_reads(w, x, y, device=dst.device), dims=[1])
        view[x - a:y - a, w["L"]] = 0


def c_generator():
    """oracle/_build/libgenreads.so (gen_reads.c): the same generators in C, for feeding the
    reference binary quickly.  Test/bench infrastructure; returns None when it is not built."""
    import ctypes
    import os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "libgenreads.so")
    if not os.path.exists(p):
        return None
    lib = ctypes.CDLL(p)
    lib.gen_reads.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                              ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.gen_reads.restype = None
    return lib


def c_reads(lib, w: dict, a: int, b: int, lines: bool, threads: int = 2) -> np.ndarray:
    """Reads a..b-1 through the C generator: (b-a, L+1) uint8, text lines or nt6 codes + NUL."""
    out = np.empty((b - a, w["L"] + 1), dtype=np.uint8)
    g = w["kind"] == "G"
    lib.gen_reads(1 if g else 0, a, b, w["L"], w["seed"], w.get("genome_seed", w["seed"]), _glen(w) if g else 0,
                  int(w["err"] * (1 << 24)) if g else 0, out.ctypes.data, 1 if lines else 0, threads)
    return out
