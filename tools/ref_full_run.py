#!/usr/bin/env python
"""Run the UNMODIFIED reference binary (oracle/_ref/ropebwt2) on a full-size seeded workload and
record {md5 of its text output, hot-path seconds, wall seconds} in tests/golden/ref_full_runs.json.

TEST / BENCH INFRASTRUCTURE, not product code.  The reads are streamed into the reference's stdin
(`ropebwt2 -LRs -`, main.c:173) chunk by chunk from the same seeded generators bench.py uses
(ropebwt2_b200/synth.py: stream_reads), so nothing of the 10 - 120 GB input ever hits the disk; the
BWT text on stdout goes through a streaming md5.  CPU only: run it in this container (or on any
box), then commit the JSON; bench.py and the -m gpu tests compare the md5 of the GPU index to it.

    python tools/ref_full_run.py --workload cfg2          # 100 M x 101 bp uniform, seed 2, -LRs
    python tools/ref_full_run.py --workload cfg3          # 1.2 B x 101 bp 30x genome, seed 3, -LRs
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ropebwt2_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_full_runs.json")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, help="a key of synth.WORKLOADS, or kind:n:L:seed")
    ap.add_argument("--flags", default="-LRs")
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads (scaled-down runs)")
    ap.add_argument("--batch", default="", help="reference -m value (default: the reference's own)")
    args = ap.parse_args()
    w = synth.workload(args.workload, args.reads)
    key = synth.workload_key(w, args.flags)
    cmd = [os.path.join(ROOT, "oracle", "_ref", "ropebwt2"), args.flags]
    if args.batch:
        cmd += ["-m", args.batch]
    cmd += ["-"]
    t0 = time.time()
    p = subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, bufsize=0)
    md5 = hashlib.md5()
    nout = [0]
    err = []

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
    for lines in synth.stream_lines(w):
        p.stdin.write(lines)
    p.stdin.close()
    to.join(), te.join()
    rc = p.wait()
    wall = time.time() - t0
    stderr = err[0].decode()
    if rc != 0:
        raise SystemExit("reference failed: " + stderr[-500:])
    hot = [float(ln.split(" symbols in ")[1].split(" sec")[0]) for ln in stderr.splitlines() if "] inserted " in ln]
    rec = {"workload": w, "flags": args.flags, "batch": args.batch or "default (-m 10415295693 bytes, main.c:94)",
           "md5_text": md5.hexdigest(), "text_bytes": nout[0], "hot_path_s": sum(hot), "hot_path_s_per_batch": hot,
           "wall_s": wall, "host_cores": os.cpu_count(), "threads": "4 workers + master",
           "gbp_per_s_hot_path": w["n"] * w["L"] / sum(hot) / 1e9, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    db[key] = rec
    json.dump(db, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
