"""The sharded build over NCCL: the library's own communicator (dlopen'ed libnccl, one process per
GPU).  World size 1 runs in-process on any GPU box; world size 2 needs two GPUs and is launched
through torchrun exactly as bench.py is."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as orc
from ropebwt2_b200 import load
from ropebwt2_b200.binding import ShardedEngine, nccl_unique_id
from ropebwt2_b200.synth import encode_batch, uniform_reads

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_communicator_world1():
    for so in (0, 1):
        e = ShardedEngine(0, so, 0, 1, nccl_uid=nccl_unique_id())
        o = orc.Oracle(so)
        for seed in (1, 2):
            buf = encode_batch(uniform_reads(4000, 50, seed, n_frac=0.01))
            e.insert_multi(buf)
            o.insert_multi(buf)
        blocks = np.concatenate([e.fetch_subbucket(s) for s in range(36)])
        assert np.array_equal(orc.decode_blocks(blocks, e.total()), o.text())
        e.close()


def test_nccl_two_ranks_torchrun():
    if load().rb2_device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tools", "shard_nccl_check.py")],
                       capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-exact") == 6, r.stdout[-2000:]
    assert r.stdout.count("dense batches 2") == 3, r.stdout[-2000:]
    if os.environ.get("RB2_P2P", "1") != "0":
        assert r.stdout.count("direct delivery 2") == 3, r.stdout[-2000:]  # the peer mappings were really used
