"""Multi-GPU plumbing for the replica / read-shard layout (one process per GPU, torch.distributed).

Round 1 runs one independent engine per rank over a disjoint shard of the reads (weak scaling,
no collective on the data path).  The only cross-rank steps are the timing
barrier and the reductions that turn per-rank measurements into one whole-job number:
units are summed, time is the MAX over ranks.  Backend-agnostic (nccl on GPUs, gloo in the CPU
tests)."""
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
