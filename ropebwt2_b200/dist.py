"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The data path of the sharded build (one index over all ranks) lives in the C
library and talks NCCL directly; torch.distributed is only used here to hand the NCCL unique id to
every rank, for the timing barrier, for the reductions that turn per-rank measurements into one
whole-job number (units are summed, time is the MAX over ranks) and to collect the sub-buckets on
rank 0 when a test wants the whole BWT.  Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_seed(base_seed: int, rank: int) -> int:
    """Generator seed of rank `rank`'s read shard: disjoint, reproducible shards without any exchange."""
    return base_seed + 1000 * rank


class Reducer:
    """all_reduce helpers on whatever backend/device the default process group uses."""

    def __init__(self, world: int, device=None):
        self.world = world
        self.device = device
        if world > 1:
            import torch.distributed as dist
            assert dist.is_initialized()
            self.dist = dist

    def _reduce(self, x: float, op_name: str) -> float:
        if self.world == 1:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op_name))
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, "MAX")

    def sum(self, x: float) -> float:
        return self._reduce(x, "SUM")

    def barrier(self) -> None:
        if self.world > 1:
            self.dist.barrier()


def whole_job_throughput(units_this_rank: float, ms_this_rank: float, red: Reducer):
    """(total units over all ranks, max time over ranks in ms, units per second)."""
    units = red.sum(units_this_rank)
    ms = red.max(ms_this_rank)
    return units, ms, (units / (ms * 1e-3) if ms > 0 else 0.0)


def broadcast_bytes(payload, src: int = 0) -> bytes:
    """Hand `payload` (bytes on rank `src`, ignored elsewhere) to every rank; identity without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return payload
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def owner_map(world: int):
    """Owner rank of each of the 36 sub-buckets (x*6+y) of a sharded index; the library's own map."""
    from .binding import load
    L = load()
    return [L.rb2_shard_owner(world, s) for s in range(36)]


def split_batch_bytes(n_strings: int, world: int):
    """[start, end) string ranges of the ranks' shares of a batch: contiguous, in rank order, so that
    (rank, position) order equals the input order (what input-order builds are defined by)."""
    cuts = [n_strings * r // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_index_blocks(my_blocks: dict, world: int, rank: int, dst: int = 0):
    """Collect {sub-bucket: uint8[n,512]} from every rank on `dst` and concatenate in sub-bucket order
    (= BWT order).  Returns the [N,512] array on `dst`, None elsewhere."""
    import numpy as np
    import torch.distributed as dist
    if world == 1:
        parts = [my_blocks]
    else:
        parts = [None] * world if rank == dst else None
        dist.gather_object(my_blocks, parts, dst=dst)
        if rank != dst:
            return None
    own = owner_map(world)
    out = []
    for s in range(36):
        for r in range(world):
            if r != own[s]:
                assert s not in parts[r] or len(parts[r][s]) == 0, f"rank {r} holds blocks of sub-bucket {s} owned by {own[s]}"
        out.append(parts[own[s]][s])
    return np.concatenate(out)
