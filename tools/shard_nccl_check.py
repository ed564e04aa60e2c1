"""Run under torchrun (one process per GPU): build ONE sharded index over all ranks through NCCL and
compare the whole BWT with the CPU oracle on rank 0.  Exit code 0 = bit-exact.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/shard_nccl_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from ropebwt2_b200.binding import ShardedEngine, nccl_unique_id
    from ropebwt2_b200.dist import broadcast_bytes, gather_index_blocks, rank_info, split_batch_bytes
    from ropebwt2_b200.synth import encode_batch, genome_reads, uniform_reads
    rank, world, local = rank_info()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    bad = 0
    # each order twice: regime by cost estimate (sparse at this size), then the dense regime forced -- whose merge
    # kernels store the new interval starts straight into the peer's memory (CUDA IPC mappings, RB2_P2P)
    for so, dense in ((1, 0), (0, 0), (2, 0), (1, 1), (0, 1), (2, 1)):
        if dense:
            os.environ["RB2_FLAT"] = "1"
        else:
            os.environ.pop("RB2_FLAT", None)
        uid = broadcast_bytes(nccl_unique_id() if rank == 0 else None)
        e = ShardedEngine(local, so, rank, world, nccl_uid=uid)
        batches = [uniform_reads(n, 60, 5 + so, n_frac=0.005), genome_reads(n // 2, 45, 6, coverage=40.0)]
        o = orc.Oracle(so) if rank == 0 else None
        for rd in batches:
            a, b = split_batch_bytes(len(rd), world)[rank]
            e.insert_multi(encode_batch(rd[a:b]))
            if rank == 0:
                o.insert_multi(encode_batch(rd))
        blocks = gather_index_blocks({s: e.fetch_subbucket(s) for s in e.owned()}, world, rank)
        if rank == 0:
            got = orc.decode_blocks(blocks, e.total())
            ok = np.array_equal(got, o.text()) and np.array_equal(e.counts(), o.counts())
            st = e.stats()
            print(f"so={so} world={world}: {e.total()} symbols {'bit-exact' if ok else 'MISMATCH'}; "
                  f"exchange {st['ms_exchange']:.1f} ms of {st['ms_total']:.1f} ms, {st['exch_bytes']} bytes received, "
                  f"dense batches {st['flat_batches']}, {'direct delivery' if st['p2p_batches'] else 'send/recv'} {st['p2p_batches']}", flush=True)
            bad += not ok
        e.quiesce()
        e.close()
    flag = torch.tensor([bad], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()) != 0)


if __name__ == "__main__":
    main()
