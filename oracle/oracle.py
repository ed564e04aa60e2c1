"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-ends for the checkers under ``oracle/``:

* ``Oracle``   -- our plain-C restatement (``bcr_oracle.c`` -> ``_build/liboracle.so``)
* ``RefLib``   -- the UNMODIFIED reference hot path compiled from ``/root/reference``
                  (``_ref/libref.so``: rle.c + rope.c + mrope.c), driven through the
                  reference's own ``mrope.h`` entry points
* ``ref_cli``  -- the UNMODIFIED reference binary (``_ref/ropebwt2``)
* ``decode_index`` -- decode any library that exports ``mr_itr_first`` /
                  ``mr_itr_next_block`` (reference or ours) into nt6 text via ``itr_text.c``

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl
reference legs may import this module.  Nothing here may read ``/root/reference`` at
run time: the ``_ref`` artefacts are prebuilt by ``oracle/Makefile`` and travel with the
repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF = os.path.join(HERE, "_ref")

_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)


def build(verbose: bool = False) -> None:
    """Compile the C restatement + helpers, and (only if the reference tree is mounted,
    i.e. in the build container) the reference artefacts under ``_ref``."""
    r = subprocess.run(["make", "-C", HERE, "oracle", "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _need(path: str) -> str:
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return path


def have_ref() -> bool:
    return os.path.exists(os.path.join(REF, "libref.so")) and os.path.exists(os.path.join(REF, "ropebwt2"))


def _as_u8(buf) -> np.ndarray:
    a = np.ascontiguousarray(buf, dtype=np.uint8)
    assert a.ndim == 1 and a.size > 0 and a[-1] == 0, "batch must end with a NUL (mrope.c:268)"
    return a


class Oracle:
    """The C restatement.  Same call shape as the multi-rope API it checks."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(_need(os.path.join(BUILD, "liboracle.so")))
            L.orc_init.restype = C.c_void_p
            L.orc_init.argtypes = [C.c_int]
            L.orc_destroy.argtypes = [C.c_void_p]
            L.orc_insert_multi.argtypes = [C.c_void_p, C.c_int64, _u8p]
            L.orc_total.restype = C.c_int64
            L.orc_total.argtypes = [C.c_void_p]
            L.orc_counts.argtypes = [C.c_void_p, _i64p]
            L.orc_text.argtypes = [C.c_void_p, _u8p]
            L.orc_rank1a.argtypes = [C.c_void_p, C.c_int64, _i64p]
            cls._lib = L
        return cls._lib

    def __init__(self, so: int = 0):
        self.h = self.lib().orc_init(so)

    def insert_multi(self, buf) -> None:
        a = _as_u8(buf)
        self.lib().orc_insert_multi(self.h, a.size, a.ctypes.data_as(_u8p))

    def total(self) -> int:
        return self.lib().orc_total(self.h)

    def counts(self) -> np.ndarray:
        c = np.zeros(36, dtype=np.int64)
        self.lib().orc_counts(self.h, c.ctypes.data_as(_i64p))
        return c.reshape(6, 6)

    def text(self) -> np.ndarray:
        out = np.empty(self.total(), dtype=np.uint8)
        self.lib().orc_text(self.h, out.ctypes.data_as(_u8p))
        return out

    def rank1a(self, x: int) -> np.ndarray:
        c = np.zeros(6, dtype=np.int64)
        self.lib().orc_rank1a(self.h, x, c.ctypes.data_as(_i64p))
        return c

    def close(self):
        if self.h:
            self.lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ITR_FIRST = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)
_ITR_NEXT = C.CFUNCTYPE(C.c_void_p, C.c_void_p)


def _itrlib():
    L = C.CDLL(_need(os.path.join(BUILD, "libitrtext.so")))
    L.itr_text.restype = C.c_int64
    L.itr_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _u8p, C.c_int64, C.c_int, _i64p, _i64p]
    L.itr_runs.restype = C.c_int64
    L.itr_runs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _u8p, _i64p, C.c_int64]
    L.blocks_text.restype = C.c_int64
    L.blocks_text.argtypes = [_u8p, C.c_int64, _u8p, C.c_int64]
    return L


def decode_blocks(blocks: np.ndarray, total: int) -> np.ndarray:
    """Decode an [n, 512] uint8 array of leaf blocks (rle.h layout) into ``total`` nt6 symbols."""
    il = _itrlib()
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    out = np.empty(max(total, 1), dtype=np.uint8)
    n = il.blocks_text(blocks.ctypes.data_as(_u8p), blocks.shape[0], out.ctypes.data_as(_u8p), total)
    if n != total:
        raise AssertionError(f"blocks decode to {n} symbols, expected {total}")
    return out[:total]


def blocks_ascii(blocks: np.ndarray, cap: int) -> np.ndarray:
    """Decode an [n, 512] uint8 array of leaf blocks into the characters the reference prints (at most cap)."""
    il = _itrlib()
    il.blocks_text2.restype = C.c_int64
    il.blocks_text2.argtypes = [_u8p, C.c_int64, _u8p, C.c_int64, C.c_int]
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    n = il.blocks_text2(blocks.ctypes.data_as(_u8p), blocks.shape[0], out.ctypes.data_as(_u8p), cap, 1)
    if n < 0:
        raise AssertionError(f"blocks decode to {-n} symbols, more than {cap}")
    return out[:n]


def decode_index(lib: C.CDLL, mr, total: int, to_free: int = 0, ascii: bool = False):
    """Decode the index behind ``mr`` (an ``mrope_t*`` of ``lib``) into ``total`` symbols by
    walking ``lib``'s own mr_itr_first / mr_itr_next_block.  Returns (text, n_blocks, n_runs)."""
    il = _itrlib()
    first = C.cast(lib.mr_itr_first, C.c_void_p)
    nxt = C.cast(lib.mr_itr_next_block, C.c_void_p)
    out = np.empty(max(total, 1), dtype=np.uint8)
    nb, nr = C.c_int64(0), C.c_int64(0)
    n = il.itr_text(mr, first, nxt, to_free, out.ctypes.data_as(_u8p), total, int(ascii), C.byref(nb), C.byref(nr))
    if n != total:
        raise AssertionError(f"index decodes to {n} symbols, expected {total}")
    return out[:total], nb.value, nr.value


def index_md5(lib: C.CDLL, mr, to_free: int = 0, chunk: int = 1 << 26):
    """md5 of the text `ropebwt2 -LR...` would print for the index behind ``mr`` (main.c:308-323:
    the BWT as "$ACGTN" characters + newline), streamed block by block through ``lib``'s own
    iterator.  Returns (hex digest, number of symbols)."""
    import hashlib
    il = _itrlib()
    il.itr_stream_open.restype = C.c_void_p
    il.itr_stream_open.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    il.itr_stream_read.restype = C.c_int64
    il.itr_stream_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    il.itr_stream_close.argtypes = [C.c_void_p]
    st = il.itr_stream_open(mr, C.cast(lib.mr_itr_first, C.c_void_p), C.cast(lib.mr_itr_next_block, C.c_void_p), to_free)
    buf = np.empty(chunk, dtype=np.uint8)
    h, total = hashlib.md5(), 0
    while True:
        n = il.itr_stream_read(st, buf.ctypes.data, chunk, 1)
        if n <= 0:
            break
        h.update(memoryview(buf[:n]))
        total += n
    il.itr_stream_close(st)
    h.update(b"\n")
    return h.hexdigest(), total


def ref_stream_run(w: dict, flags: str = "-LRs", batch: str = "", want_md5: bool = True, gen_threads: int = 2) -> dict:
    """Run the UNMODIFIED reference binary on workload ``w`` (ropebwt2_b200.synth.workload): the reads
    are streamed into its stdin (``ropebwt2 <flags> -``, main.c:173) from the seeded C generator, its
    text output goes through a streaming md5 (or to /dev/null).  Returns md5, the reference's own
    hot-path timer (sum of the ``inserted ... in X sec`` lines, main.c:241,249) and wall seconds."""
    import hashlib
    import threading
    import time
    from ropebwt2_b200 import synth
    cmd = [_need(os.path.join(REF, "ropebwt2")), flags]
    if batch:
        cmd += ["-m", batch]
    if not want_md5:
        cmd += ["-o", "/dev/null"]
    cmd += ["-"]
    t0 = time.time()
    p = subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, bufsize=0)
    md5, nout, err = hashlib.md5(), [0], []

    def pump_out():
        while True:
            b = p.stdout.read(1 << 24)
            if not b:
                break
            md5.update(b)
            nout[0] += len(b)

    def pump_err():
        err.append(p.stderr.read())

    to, te = threading.Thread(target=pump_out), threading.Thread(target=pump_err)
    to.start(), te.start()
    try:
        for lines in synth.stream_lines(w, threads=gen_threads):
            mv = memoryview(lines)
            while len(mv):  # (an unbuffered pipe write may be partial: one write() moves at most 2^31 - 4096 bytes)
                mv = mv[p.stdin.write(mv):]
        p.stdin.close()
    except BrokenPipeError:
        pass
    to.join(), te.join()
    rc = p.wait()
    wall = time.time() - t0
    stderr = err[0].decode()
    if rc != 0:
        raise RuntimeError("reference failed: " + stderr[-500:])
    hot = [float(ln.split(" symbols in ")[1].split(" sec")[0]) for ln in stderr.splitlines() if "] inserted " in ln]
    if nout[0] and nout[0] != w["n"] * (w["L"] + 1) * (2 if "R" not in flags else 1) + 1:
        raise RuntimeError(f"reference printed {nout[0]} symbols for {w['n']} reads of {w['L']}")
    return {"workload": w, "flags": flags, "batch": batch or "default (-m 10415295693 bytes, main.c:94)",
            "md5_text": md5.hexdigest() if want_md5 else None, "text_bytes": nout[0], "hot_path_s": sum(hot), "hot_path_s_per_batch": hot,
            "wall_s": wall, "host_cores": os.cpu_count(), "threads": "4 workers + master",
            "gbp_per_s_hot_path": w["n"] * w["L"] / sum(hot) / 1e9, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}


GOLDEN_RUNS = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_full_runs.json")


def ref_recorded(w: dict, flags: str):
    """The recorded full run of the reference for this workload (tools/ref_full_run.py), or None."""
    import json
    from ropebwt2_b200 import synth
    if not os.path.exists(GOLDEN_RUNS):
        return None
    return json.load(open(GOLDEN_RUNS)).get(synth.workload_key(w, flags))


class RefLib:
    """The unmodified reference ``mrope.h`` API (``oracle/_ref/libref.so``)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(_need(os.path.join(REF, "libref.so")))
            L.mr_init.restype = C.c_void_p
            L.mr_init.argtypes = [C.c_int, C.c_int, C.c_int]
            L.mr_destroy.argtypes = [C.c_void_p]
            L.mr_insert_multi.argtypes = [C.c_void_p, C.c_int64, _u8p, C.c_int]
            L.mr_insert1.restype = C.c_int64
            L.mr_insert1.argtypes = [C.c_void_p, _u8p]
            L.mr_rank2a.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _i64p, _i64p]
            L.mr_itr_first.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            L.mr_itr_next_block.restype = C.c_void_p
            L.mr_itr_next_block.argtypes = [C.c_void_p]
            L.mr_thr_min.restype = C.c_int
            L.mr_thr_min.argtypes = [C.c_void_p, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, so: int = 0, max_nodes: int = 64, block_len: int = 512):
        self.h = self.lib().mr_init(max_nodes, block_len, so)
        self.n_sym = 0

    def insert_multi(self, buf, is_thr: int = 0) -> None:
        a = _as_u8(buf)
        self.lib().mr_insert_multi(self.h, a.size, a.ctypes.data_as(_u8p), is_thr)
        self.n_sym += a.size

    def insert1(self, s) -> None:
        a = np.ascontiguousarray(s, dtype=np.uint8)
        assert a[-1] == 0
        self.lib().mr_insert1(self.h, a.ctypes.data_as(_u8p))
        self.n_sym += a.size

    def total(self) -> int:
        return self.n_sym

    def text(self) -> np.ndarray:
        return decode_index(self.lib(), self.h, self.n_sym)[0]

    def rank2a(self, x: int, y: int):
        cx = np.zeros(6, dtype=np.int64)
        cy = np.zeros(6, dtype=np.int64)
        self.lib().mr_rank2a(self.h, x, y, cx.ctypes.data_as(_i64p), cy.ctypes.data_as(_i64p) if y >= 0 else None)
        return cx, cy

    def close(self):
        if self.h:
            self.lib().mr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_cli(args, stdin_bytes: bytes = b"", timeout: float = 3600.0):
    """Run the unmodified reference binary; returns (stdout bytes, stderr text)."""
    exe = _need(os.path.join(REF, "ropebwt2"))
    r = subprocess.run([exe] + list(args), input=stdin_bytes, capture_output=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ropebwt2 {args} failed: {r.stderr.decode()[-400:]}")
    return r.stdout, r.stderr.decode()


def ref_hot_path_seconds(stderr_text: str) -> float:
    """Sum of the reference's own hot-path timer lines
    ``[M::main_ropebwt2] inserted N symbols in X sec`` (main.c:241,249)."""
    tot = 0.0
    for line in stderr_text.splitlines():
        if "inserted" in line and " sec," in line:
            tot += float(line.split(" in ")[1].split(" sec")[0])
    return tot
