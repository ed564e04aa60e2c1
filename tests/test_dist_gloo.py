"""CPU test of the N>1 host path: world_size-2 gloo process group, the same reductions bench.py uses
to turn per-rank replica measurements into one whole-job number (sum of units / max of time), and the
disjoint read shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ropebwt2_b200.dist import Reducer, rank_info, shard_seed, whole_job_throughput
from ropebwt2_b200.synth import uniform_reads


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, w, l = rank_info()
        assert (r, w, l) == (rank, world, rank)
        red = Reducer(w)
        red.barrier()
        # every rank "measures" a different time for the same amount of work
        units, ms, thr = whole_job_throughput(1000.0, 10.0 * (rank + 1), red)
        reads = uniform_reads(64, 20, shard_seed(7, r))
        digest = float(np.frombuffer(reads.tobytes(), dtype=np.uint8).astype(np.int64).sum())
        out.put((rank, units, ms, thr, red.max(digest), red.sum(digest), digest))
    finally:
        dist.destroy_process_group()


def test_world_size_2_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, units, ms, thr, dmax, dsum, digest in res:
        assert units == 2000.0                 # units are summed over ranks
        assert ms == 20.0                      # time is the max over ranks
        assert abs(thr - 2000.0 / 0.020) < 1e-6
    # shards differ between ranks (disjoint seeds) and the reductions saw both
    assert res[0][6] != res[1][6]
    assert res[0][5] == res[0][6] + res[1][6]
    assert res[0][4] == max(res[0][6], res[1][6])


def test_single_process_defaults():
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    assert rank_info() == (0, 1, 0)
    red = Reducer(1)
    assert whole_job_throughput(5.0, 2.0, red) == (5.0, 2.0, 2500.0)


# ---- host logic of the sharded build (one index over all ranks) -----------------------------------

def _shard_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ropebwt2_b200.dist import broadcast_bytes, gather_index_blocks, owner_map, split_batch_bytes
        # the NCCL unique id travels from rank 0 to everybody (128 opaque bytes)
        uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None)
        own = owner_map(world)
        # every rank contributes fake leaf blocks for the sub-buckets it owns: block byte 0 = sub-bucket id
        mine = {s: np.full((2 if s % 5 == 0 else 1, 512), s, dtype=np.uint8) for s in range(36) if own[s] == rank}
        blocks = gather_index_blocks(mine, world, rank)
        out.put((rank, uid, split_batch_bytes(1001, world)[rank], None if blocks is None else blocks[:, 0].tolist()))
    finally:
        dist.destroy_process_group()


def test_world_size_2_sharded_host_logic():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == bytes(range(128))
    # contiguous shares in rank order that cover the batch
    assert res[0][2][0] == 0 and res[0][2][1] == res[1][2][0] and res[1][2][1] == 1001
    # rank 0 sees all sub-buckets in BWT order (sub-bucket ids ascending), nothing on the other rank
    assert res[1][3] is None
    ids = res[0][3]
    assert ids == sorted(ids) and sorted(set(ids)) == list(range(36))


def test_owner_map_is_a_partition_into_contiguous_ranges():
    from ropebwt2_b200.dist import owner_map
    for world in range(1, 9):
        own = owner_map(world)
        assert len(own) == 36 and own == sorted(own) and set(own) == set(range(world))
        # the 16 ACGT x ACGT sub-buckets (the ones that carry the data) are spread evenly
        main = [own[x * 6 + y] for x in range(1, 5) for y in range(1, 5)]
        counts = [main.count(r) for r in range(world)]
        assert max(counts) - min(counts) <= 1
    from ropebwt2_b200 import load
    assert load().rb2_shard_owner(9, 0) == -1 and load().rb2_shard_owner(2, 36) == -1


# ---- the sharded algorithm itself, modelled on the CPU (oracle/shard_model.py) -------------------------

def _batches(so):
    from ropebwt2_b200.synth import genome_reads, varlen_reads
    rd = genome_reads(160, 14, 31 + so, coverage=25.0)       # duplicates -> non-empty intervals in later batches
    return [list(rd[:70]), list(rd[70:110]), varlen_reads(40, 12, 5 + so)]


def _reversed(reads):
    return [[int(c) for c in r[::-1]] for r in reads]


def _oracle_text(so, batches):
    from oracle import oracle as orc
    from ropebwt2_b200.synth import encode_batch
    o = orc.Oracle(so)
    for b in batches:
        o.insert_multi(encode_batch(b))
    return o.text().tolist()


def test_shard_model_world1_matches_the_oracle():
    from oracle.shard_model import LocalComm, ShardModel, owner_map
    from ropebwt2_b200 import load
    for world in range(1, 9):   # the model partitions exactly like the library
        assert owner_map(world) == [load().rb2_shard_owner(world, s) for s in range(36)]
    for so in (0, 1, 2):
        batches = _batches(so)
        m = ShardModel(LocalComm(), so)
        for b in batches:
            m.insert_multi(_reversed(b))
        assert m.text() == _oracle_text(so, batches), so


def _model_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.shard_model import ShardModel
        from ropebwt2_b200.dist import split_batch_bytes

        class GlooComm:
            def __init__(self):
                self.rank, self.world = rank, world

            def allgather(self, obj):
                box = [None] * world
                dist.all_gather_object(box, obj)
                return box
        ok = True
        for so in (0, 1, 2):
            batches = _batches(so)
            m = ShardModel(GlooComm(), so)
            for b in batches:
                a, e = split_batch_bytes(len(b), world)[rank]
                m.insert_multi(_reversed(b[a:e]))
            text = m.text()
            # every rank holds only its own sub-buckets, and together they are the oracle's BWT
            assert set(m.seg) == {s for s in range(36) if m.own[s] == rank}
            if rank == 0:
                ok = ok and text == _oracle_text(so, batches)
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shard_model_over_gloo_matches_the_oracle(world):
    """The N > 1 algorithm on CPU: sub-bucket ownership, per-column table all-gather, whole-index ranks
    from local counts + gathered totals, routing (x,y)+a -> (a,x), target order -- over a real process
    group.  The CUDA engine implements the same steps (tests/test_sharded_gpu.py checks it on the GPU)."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_model_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
