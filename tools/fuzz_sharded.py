#!/usr/bin/env python
"""Random scenarios for the sharded build against the CPU oracle (TEST INFRASTRUCTURE): ranks, sorting order, regime
per batch, direct delivery on/off per batch, batch sizes including empty shares and length outliers, block fetches and
resets in between.  Runs on a GPU, or on the CPU emulator of the kernels:

    RB2_EMU=1 python tools/fuzz_sharded.py [--seed S] [--rounds N]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rounds", type=int, default=20)
    args = ap.parse_args()
    if os.environ.get("RB2_EMU") == "1":
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        from ropebwt2_b200 import binding
        binding.load(path=build_emu.build())
    from oracle import oracle as orc
    from ropebwt2_b200.synth import encode_batch
    from test_sharded_gpu import Cluster, split
    rng = np.random.default_rng(args.seed)
    for it in range(args.rounds):
        so, P = int(rng.integers(0, 3)), int(rng.choice([2, 3, 4, 8]))
        o, c = orc.Oracle(so), Cluster(so, P)
        log = [f"so={so} P={P}"]
        for b in range(int(rng.integers(1, 6))):
            kind = rng.integers(0, 4)
            n = int(rng.integers(1, 1500))
            if kind == 0:
                strs = [rng.integers(1, 5, size=int(rng.integers(20, 40))).astype(np.uint8) for _ in range(n)]
            elif kind == 1:
                L = int(rng.integers(1, 60))
                strs = [rng.integers(1, 6, size=L).astype(np.uint8) for _ in range(n)]
            elif kind == 2:
                strs = [rng.integers(1, 5, size=int(rng.integers(0, 30))).astype(np.uint8) for _ in range(n // 4 + 1)]
                strs.insert(int(rng.integers(0, len(strs) + 1)), rng.integers(1, 5, size=int(rng.integers(200, 1500))).astype(np.uint8))
            else:
                base = [rng.integers(1, 5, size=12).astype(np.uint8) for _ in range(3)]
                strs = [base[int(rng.integers(0, 3))].copy() for _ in range(n)]
            flat, p2p = str(int(rng.integers(0, 2))), str(int(rng.integers(0, 2)))
            os.environ["RB2_FLAT"], os.environ["RB2_P2P"] = flat, p2p
            os.environ["RB2_SPLIT_SLACK"] = str(int(rng.choice([4096, 64 << 20])))
            rev = bool(rng.integers(0, 2))
            log.append(f"batch {b}: kind {kind} n {len(strs)} flat {flat} p2p {p2p} rev {rev} slack {os.environ['RB2_SPLIT_SLACK']}")
            o.insert_multi(encode_batch(strs, True, rev))
            if rng.integers(0, 4) == 0:  # everything on one rank
                parts = [[] for _ in range(P)]
                parts[int(rng.integers(0, P))] = strs
            else:
                parts = split(strs, P)
            c.insert([encode_batch(p, True, rev) if len(p) else np.zeros(0, np.uint8) for p in parts])
            if rng.integers(0, 3) == 0:
                assert np.array_equal(c.text(), o.text()), "\n".join(log)
        ok = np.array_equal(c.text(), o.text()) and all(np.array_equal(e.counts(), o.counts()) for e in c.eng)
        print(("ok   " if ok else "FAIL ") + " | ".join(log), flush=True)
        assert ok
        c.close()


if __name__ == "__main__":
    main()
