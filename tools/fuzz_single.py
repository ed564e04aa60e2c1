#!/usr/bin/env python
"""Random scenarios for the one-GPU engine behind mrope.h against the CPU oracle (TEST INFRASTRUCTURE): sorting order,
regime per batch, fused column kernel on/off, synchronous / pipelined calls, slice size, batch shapes (uniform, ragged,
length outliers, duplicates, empty strings), mr_insert1, rank queries, dump / restore and text checks in between.

    RB2_EMU=1 python tools/fuzz_single.py [--seed S] [--rounds N]      (CPU emulator of the kernels)
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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
    from ropebwt2_b200 import MRope, load
    from ropebwt2_b200.synth import encode_batch
    rng = np.random.default_rng(args.seed)
    tmp = tempfile.mkdtemp()
    for it in range(args.rounds):
        so = int(rng.integers(0, 3))
        os.environ["RB2_ASYNC"] = str(int(rng.integers(0, 2)))  # read when the engine is created
        os.environ["RB2_FUSED"] = str(int(rng.integers(0, 2)))
        o, m = orc.Oracle(so), MRope(so)
        log = [f"so={so} async={os.environ['RB2_ASYNC']} fused={os.environ['RB2_FUSED']}"]

        def check(what):
            got = orc.decode_index(load(), m.h, m.total())[0]
            assert np.array_equal(got, o.text()), what + "\n" + "\n".join(log)

        for b in range(int(rng.integers(1, 7))):
            kind = int(rng.integers(0, 6))
            n = int(rng.integers(1, 2500))
            if kind == 0:
                strs = [rng.integers(1, 5, size=int(rng.integers(20, 40))).astype(np.uint8) for _ in range(n)]
            elif kind == 1:
                L = int(rng.integers(1, 70))
                strs = [rng.integers(1, 6, size=L).astype(np.uint8) for _ in range(n)]
            elif kind == 2:
                strs = [rng.integers(1, 5, size=int(rng.integers(0, 30))).astype(np.uint8) for _ in range(n // 4 + 1)]
                strs.insert(int(rng.integers(0, len(strs) + 1)), rng.integers(1, 5, size=int(rng.integers(200, 2500))).astype(np.uint8))
            elif kind == 3:
                base = [rng.integers(1, 5, size=12).astype(np.uint8) for _ in range(3)]
                strs = [base[int(rng.integers(0, 3))].copy() for _ in range(n)]
            elif kind == 4:  # one string through mr_insert1
                s = rng.integers(1, 6, size=int(rng.integers(0, 300))).astype(np.uint8)
                log.append(f"step {b}: insert1 len {len(s)}")
                o.insert_multi(encode_batch([s]))
                m.insert1(encode_batch([s]))
                continue
            else:  # dump, restore into a fresh engine, go on with it
                p = os.path.join(tmp, f"f{it}_{b}.fmr")
                log.append(f"step {b}: dump + restore")
                m.dump(p)
                m.close()
                m = MRope.restore(p)
                check("after restore")
                continue
            os.environ["RB2_FLAT"] = str(int(rng.integers(0, 2)))
            os.environ["RB2_WIDE_RATIO"] = str(int(rng.choice([1, 192, 10 ** 9])))
            os.environ["RB2_SPLIT_SLACK"] = str(int(rng.choice([4096, 64 << 20])))
            rev = bool(rng.integers(0, 2))
            log.append(f"step {b}: kind {kind} n {len(strs)} flat {os.environ['RB2_FLAT']} wide {os.environ['RB2_WIDE_RATIO']} rev {rev} slack {os.environ['RB2_SPLIT_SLACK']}")
            buf = encode_batch(strs, True, rev)
            o.insert_multi(buf)
            m.insert_multi(buf)
            r = int(rng.integers(0, 4))
            if r == 0:
                check("after batch")
            elif r == 1:
                assert np.array_equal(m.counts(), o.counts()), "\n".join(log)
            elif r == 2 and o.total() > 0:
                xs = rng.integers(0, o.total() + 1, size=5)
                for x in xs.tolist():
                    assert np.array_equal(m.rank2a(int(x))[0], o.rank1a(int(x))), "\n".join(log)
        check("at the end")
        assert np.array_equal(m.counts(), o.counts()), "\n".join(log)
        print("ok   " + " | ".join(log), flush=True)
        m.close()


if __name__ == "__main__":
    main()
