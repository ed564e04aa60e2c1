#!/usr/bin/env python
"""Run the UNMODIFIED reference binary (oracle/_ref/ropebwt2) on a full-size seeded workload and
record {md5 of its text output, hot-path seconds, wall seconds} in tests/golden/ref_full_runs.json.

TEST / BENCH INFRASTRUCTURE, not product code.  The reads are streamed into the reference's stdin
chunk by chunk from the seeded generators bench.py uses (oracle.ref_stream_run), so nothing of the
10 - 120 GB input ever hits the disk; the BWT text on stdout goes through a streaming md5.  CPU only:
run it in this container (or on any box), then commit the JSON; bench.py and the -m gpu tests compare
the md5 of the GPU index to it.

    python tools/ref_full_run.py --workload cfg2          # 100 M x 101 bp uniform, seed 2, -LRs
    python tools/ref_full_run.py --workload cfg3          # 1.2 B x 101 bp 30x genome, seed 3, -LRs
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from ropebwt2_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, help="a key of synth.WORKLOADS, or kind:n:L:seed")
    ap.add_argument("--flags", default="-LRs")
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads (scaled-down runs)")
    ap.add_argument("--batch", default="", help="reference -m value (default: the reference's own)")
    args = ap.parse_args()
    w = synth.workload(args.workload, args.reads)
    rec = orc.ref_stream_run(w, args.flags, args.batch)
    db = json.load(open(orc.GOLDEN_RUNS)) if os.path.exists(orc.GOLDEN_RUNS) else {}
    db[synth.workload_key(w, args.flags)] = rec
    json.dump(db, open(orc.GOLDEN_RUNS, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
