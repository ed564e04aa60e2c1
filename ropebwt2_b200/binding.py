"""ctypes binding of libropebwt2_b200.so: the reference-facing ``mrope.h`` API (class
``MRope``) and the engine-level C-ABI of ``include/ropebwt2_b200.h`` (class ``Engine``)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)

# every symbol the headers in include/ declare (tests check the library exports all of them)
MROPE_SYMBOLS = ["mr_init", "mr_destroy", "mr_thr_min", "mr_insert1", "mr_insert_multi", "mr_rank2a",
                 "mr_itr_first", "mr_itr_next_block", "mr_print_tree", "mr_dump", "mr_restore"]
ROPE_SYMBOLS = ["rope_init", "rope_destroy", "rope_insert_run", "rope_rank2a", "rope_itr_first",
                "rope_itr_next_block", "rope_print_node", "rope_dump", "rope_restore"]
RLE_SYMBOLS = ["rle_count", "rle_print"]
RB2_SYMBOLS = ["rb2_device_count", "rb2_create", "rb2_create_auto", "rb2_destroy", "rb2_sorting_order", "rb2_insert_multi",
               "rb2_insert_multi_dev", "rb2_counts", "rb2_rank2a", "rb2_num_blocks", "rb2_fetch_blocks",
               "rb2_load_blocks", "rb2_get_stats", "rb2_reset_stats", "rb2_stream", "rb2_dev_alloc",
               "rb2_dev_free", "rb2_dev_upload", "rb2_reset", "rb2_host_alloc", "rb2_host_free", "rb2_insert_run",
               "rb2_bucket_rank2a", "rb2_last_sentinel_rank",
               "rb2_group_create", "rb2_group_destroy", "rb2_nccl_unique_id", "rb2_create_sharded",
               "rb2_insert_multi_sharded", "rb2_insert_multi_sharded_dev", "rb2_sharded_quiesce", "rb2_shard_owner", "rb2_num_buckets",
               "rb2_rank_batch", "rb2_sync", "rb2_span_begin", "rb2_span_ms", "rb2_job_history"]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_strings", "n_symbols", "n_columns", "n_launches", "n_merge_launches",
                                         "merge_blocks", "merge_bytes_rw", "n_records", "pool_blocks", "pool_capacity")] + \
               [(n, C.c_double) for n in ("ms_total", "ms_h2d", "ms_transpose", "ms_members", "ms_groups", "ms_merge", "ms_directory",
                                          "ms_merge_general")] + [("general_items", C.c_int64), ("ms_exchange", C.c_double),
                                                                  ("exch_bytes", C.c_int64), ("ms_convert", C.c_double), ("flat_batches", C.c_int64),
                                                                  ("p2p_batches", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class _MRopeStruct(C.Structure):  # mrope_t, include/mrope.h
    _fields_ = [("so", C.c_uint8), ("thr_min", C.c_int), ("r", C.c_void_p * 6), ("priv", C.c_void_p)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(rebuild: bool = False, path: str = None) -> C.CDLL:
    """Build (if stale) and load the shared library.  Raises if it cannot be built/loaded:
    there is deliberately no fallback.  `path` is for the test suite only: tests/emu loads the
    engine compiled against its CPU emulator of the CUDA execution model that way (kernel-logic
    checks on machines without a GPU); nothing in this package ever passes it."""
    global _lib
    if _lib is not None and not rebuild and path is None:
        return _lib
    if path is None:
        # build() returns at once when the library is newer than every source; a prebuilt library
        # on a machine without nvcc (the GPU box) is used as it is
        import shutil
        have_nvcc = shutil.which(os.environ.get("NVCC", "nvcc")) is not None
        path = _build.build(force=rebuild) if (have_nvcc or not os.path.exists(_build.LIB)) else _build.LIB
    L = C.CDLL(path)
    L.mr_init.restype = C.c_void_p
    L.mr_init.argtypes = [C.c_int, C.c_int, C.c_int]
    L.mr_destroy.argtypes = [C.c_void_p]
    L.mr_thr_min.restype = C.c_int
    L.mr_thr_min.argtypes = [C.c_void_p, C.c_int]
    L.mr_insert1.restype = C.c_int64
    L.mr_insert1.argtypes = [C.c_void_p, _u8p]
    L.mr_insert_multi.argtypes = [C.c_void_p, C.c_int64, _u8p, C.c_int]
    L.mr_rank2a.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _i64p, _i64p]
    L.mr_itr_first.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.mr_itr_next_block.restype = C.c_void_p
    L.mr_itr_next_block.argtypes = [C.c_void_p]
    L.mr_dump.argtypes = [C.c_void_p, C.c_void_p]
    L.mr_restore.restype = C.c_void_p
    L.mr_restore.argtypes = [C.c_void_p]
    L.mr_print_tree.argtypes = [C.c_void_p]
    L.rope_init.restype = C.c_void_p
    L.rope_init.argtypes = [C.c_int, C.c_int]
    L.rope_destroy.argtypes = [C.c_void_p]
    L.rope_insert_run.restype = C.c_int64
    L.rope_insert_run.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p]
    L.rope_rank2a.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _i64p, _i64p]
    L.rope_itr_first.argtypes = [C.c_void_p, C.c_void_p]
    L.rope_itr_next_block.restype = C.c_void_p
    L.rope_itr_next_block.argtypes = [C.c_void_p]
    L.rb2_device_count.restype = C.c_int
    L.rb2_create.restype = C.c_void_p
    L.rb2_create.argtypes = [C.c_int, C.c_int]
    L.rb2_destroy.argtypes = [C.c_void_p]
    L.rb2_insert_multi.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.rb2_insert_multi_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.rb2_counts.argtypes = [C.c_void_p, _i64p]
    L.rb2_rank2a.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _i64p, _i64p]
    L.rb2_num_blocks.restype = C.c_int64
    L.rb2_num_blocks.argtypes = [C.c_void_p, C.c_int]
    L.rb2_fetch_blocks.restype = C.c_int64
    L.rb2_fetch_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, _u8p, _i64p]
    L.rb2_load_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int64, _u8p, _i64p]
    L.rb2_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.rb2_reset_stats.argtypes = [C.c_void_p]
    L.rb2_stream.restype = C.c_void_p
    L.rb2_stream.argtypes = [C.c_void_p]
    L.rb2_dev_alloc.restype = C.c_void_p
    L.rb2_dev_alloc.argtypes = [C.c_void_p, C.c_int64]
    L.rb2_dev_free.argtypes = [C.c_void_p, C.c_void_p]
    L.rb2_dev_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.rb2_reset.argtypes = [C.c_void_p]
    L.rb2_host_alloc.restype = C.c_void_p
    L.rb2_host_alloc.argtypes = [C.c_int64]
    L.rb2_host_free.argtypes = [C.c_void_p]
    L.rb2_insert_run.restype = C.c_int64
    L.rb2_insert_run.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64]
    L.rb2_bucket_rank2a.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, _i64p, _i64p]
    L.rb2_last_sentinel_rank.restype = C.c_int64
    L.rb2_last_sentinel_rank.argtypes = [C.c_void_p]
    L.rb2_group_create.restype = C.c_void_p
    L.rb2_group_create.argtypes = [C.c_int]
    L.rb2_group_destroy.argtypes = [C.c_void_p]
    L.rb2_nccl_unique_id.argtypes = [_u8p]
    L.rb2_create_sharded.restype = C.c_void_p
    L.rb2_create_sharded.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.rb2_insert_multi_sharded.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.rb2_insert_multi_sharded_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.rb2_sharded_quiesce.restype = None
    L.rb2_sharded_quiesce.argtypes = [C.c_void_p]
    L.rb2_shard_owner.restype = C.c_int
    L.rb2_shard_owner.argtypes = [C.c_int, C.c_int]
    L.rb2_rank_batch.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p]
    L.rb2_sync.argtypes = [C.c_void_p]
    L.rb2_span_begin.argtypes = [C.c_void_p]
    L.rb2_span_ms.restype = C.c_double
    L.rb2_span_ms.argtypes = [C.c_void_p]
    L.rb2_job_history.restype = C.c_int
    L.rb2_job_history.argtypes = [C.c_void_p, C.POINTER(Stats), C.c_int]
    L.rb2_num_buckets.restype = C.c_int
    L.rb2_num_buckets.argtypes = [C.c_void_p]
    _lib = L
    return L


def _libc():
    return C.CDLL(None)


class MRope:
    """The reference's multi-rope API (include/mrope.h) as a user of the reference would call it."""

    def __init__(self, so: int = 0, max_nodes: int = 64, block_len: int = 512, _handle=None):
        self.L = load()
        self.h = _handle if _handle is not None else self.L.mr_init(max_nodes, block_len, so)

    @property
    def struct(self) -> _MRopeStruct:
        return _MRopeStruct.from_address(self.h)

    @property
    def engine_handle(self):
        # rb2_priv_t starts with the engine pointer (csrc/mrope_b200.c)
        return C.c_void_p.from_address(self.struct.priv).value

    def insert_multi(self, buf, is_thr: int = 1) -> None:
        a = np.ascontiguousarray(buf, dtype=np.uint8)
        self.L.mr_insert_multi(self.h, a.size, a.ctypes.data_as(_u8p), is_thr)

    def insert1(self, s) -> int:
        a = np.ascontiguousarray(s, dtype=np.uint8)
        assert a[-1] == 0
        return self.L.mr_insert1(self.h, a.ctypes.data_as(_u8p))

    def counts(self) -> np.ndarray:
        """c[b][a] read the way the inline mr_get_c does: through mrope_t::r[b]->c (offset 8 in rope_t)."""
        out = np.zeros((6, 6), dtype=np.int64)
        st = self.struct
        for b in range(6):
            if st.r[b]:
                out[b] = np.ctypeslib.as_array((C.c_int64 * 6).from_address(st.r[b] + 8))
        return out

    def total(self) -> int:
        return int(self.counts().sum())

    def rank2a(self, x: int, y: int = -1):
        cx = np.zeros(6, dtype=np.int64)
        cy = np.zeros(6, dtype=np.int64)
        self.L.mr_rank2a(self.h, x, y, cx.ctypes.data_as(_i64p), cy.ctypes.data_as(_i64p) if y >= 0 else None)
        return cx, cy

    def rank_batch(self, xs) -> np.ndarray:
        """occ(a, x) for every position in xs at once -> int64 [n, 6] (rb2_rank_batch)."""
        x = np.ascontiguousarray(xs, dtype=np.int64)
        out = np.zeros((x.size, 6), dtype=np.int64)
        self.L.rb2_rank_batch(self.engine_handle, x.size, x.ctypes.data_as(_i64p), out.ctypes.data_as(_i64p))
        return out

    def stats(self) -> dict:
        st = Stats()
        self.L.rb2_get_stats(self.engine_handle, C.byref(st))
        return st.as_dict()

    def reset_stats(self) -> None:
        self.L.rb2_reset_stats(self.engine_handle)

    def job_history(self, n: int = 64):
        """Per-batch statistics of the last n mr_insert_multi calls (waits for the queued batches)."""
        arr = (Stats * n)()
        k = self.L.rb2_job_history(self.engine_handle, arr, n)
        return [arr[i].as_dict() for i in range(k)]

    def dump(self, path: str) -> None:
        libc = _libc()
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        fp = libc.fopen(path.encode(), b"wb")
        if not fp:
            raise OSError(path)
        self.L.mr_dump(self.h, fp)
        libc.fclose(fp)

    @classmethod
    def restore(cls, path: str) -> "MRope":
        L = load()
        libc = _libc()
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        fp = libc.fopen(path.encode(), b"rb")
        if not fp:
            raise OSError(path)
        h = L.mr_restore(fp)
        libc.fclose(fp)
        return cls(_handle=h)

    def close(self) -> None:
        if self.h:
            self.L.mr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """Engine-level C-ABI (include/ropebwt2_b200.h), used by bench.py for the device-resident leg."""

    def __init__(self, device: int = 0, so: int = 0):
        self.L = load()
        self.h = self.L.rb2_create(device, so)

    def insert_multi(self, buf) -> None:
        a = np.ascontiguousarray(buf, dtype=np.uint8)
        self.L.rb2_insert_multi(self.h, a.size, a.ctypes.data)

    def insert_multi_ptr(self, host_ptr: int, n: int) -> None:
        self.L.rb2_insert_multi(self.h, n, host_ptr)

    def insert_multi_dev(self, dev_ptr: int, n: int) -> None:
        self.L.rb2_insert_multi_dev(self.h, n, dev_ptr)

    def dev_alloc(self, n: int) -> int:
        return self.L.rb2_dev_alloc(self.h, n)

    def dev_free(self, p: int) -> None:
        self.L.rb2_dev_free(self.h, p)

    def dev_upload(self, dst: int, src) -> None:
        a = np.ascontiguousarray(src, dtype=np.uint8)
        self.L.rb2_dev_upload(self.h, dst, a.ctypes.data, a.size)

    def counts(self) -> np.ndarray:
        c = np.zeros(36, dtype=np.int64)
        self.L.rb2_counts(self.h, c.ctypes.data_as(_i64p))
        return c.reshape(6, 6)

    def stats(self) -> dict:
        st = Stats()
        self.L.rb2_get_stats(self.h, C.byref(st))
        return st.as_dict()

    def reset_stats(self) -> None:
        self.L.rb2_reset_stats(self.h)

    def reset(self) -> None:
        self.L.rb2_reset(self.h)

    def total(self) -> int:
        return int(self.counts().sum())

    def fetch_all_blocks(self):
        """All leaf blocks, buckets 0..5 in order -> (uint8 [n,512], int64 [n,6])."""
        blks, cnts = [], []
        for b in range(6):
            n = self.L.rb2_num_blocks(self.h, b)
            buf = np.zeros((n, 512), dtype=np.uint8)
            cnt = np.zeros((n, 6), dtype=np.int64)
            got = self.L.rb2_fetch_blocks(self.h, b, 0, n, buf.ctypes.data_as(_u8p), cnt.ctypes.data_as(_i64p))
            assert got == n
            blks.append(buf)
            cnts.append(cnt)
        return np.concatenate(blks), np.concatenate(cnts)

    def close(self) -> None:
        if self.h:
            self.L.rb2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedEngine(Engine):
    """One rank of a sharded build (one index over several GPUs, include/ropebwt2_b200.h).
    ``group``: handle from ``local_group()`` when the ranks are threads of this process;
    ``nccl_uid``: 128 bytes from ``nccl_unique_id()`` when they are processes (one per GPU)."""

    def __init__(self, device: int, so: int, rank: int, nranks: int, group=None, nccl_uid: bytes = None):
        self.L = load()
        self.rank, self.nranks = rank, nranks
        uid = (C.c_uint8 * 128).from_buffer_copy(nccl_uid) if nccl_uid is not None else None
        self.h = self.L.rb2_create_sharded(device, so, rank, nranks, group, uid)

    def insert_multi(self, buf) -> None:
        """Collective: every rank passes its own share of the batch (may be empty)."""
        a = np.ascontiguousarray(buf, dtype=np.uint8)
        self.L.rb2_insert_multi_sharded(self.h, a.size, a.ctypes.data if a.size else None)

    def insert_multi_ptr(self, host_ptr: int, n: int) -> None:
        self.L.rb2_insert_multi_sharded(self.h, n, host_ptr)

    def insert_multi_dev(self, dev_ptr: int, n: int) -> None:
        self.L.rb2_insert_multi_sharded_dev(self.h, n, dev_ptr)

    def quiesce(self) -> None:
        """Collective: close the peer mappings of the direct delivery (before ranks close at different times)."""
        self.L.rb2_sharded_quiesce(self.h)

    def owned(self):
        return [s for s in range(36) if self.L.rb2_shard_owner(self.nranks, s) == self.rank]

    def fetch_subbucket(self, s: int) -> np.ndarray:
        n = self.L.rb2_num_blocks(self.h, s)
        buf = np.zeros((n, 512), dtype=np.uint8)
        if n:
            got = self.L.rb2_fetch_blocks(self.h, s, 0, n, buf.ctypes.data_as(_u8p), None)
            assert got == n
        return buf


def local_group(nranks: int):
    return load().rb2_group_create(nranks)


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    load().rb2_nccl_unique_id(buf)
    return bytes(buf)
