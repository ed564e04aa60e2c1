"""TEST INFRASTRUCTURE.  Generates tests/golden/golden.json by running the UNMODIFIED reference
binary (oracle/_ref/ropebwt2, built from /root/reference by oracle/Makefile) on seeded synthetic
inputs.  Run in the build container:  python oracle/gen_golden.py

Each case records the generator parameters, the reference command-line flags, and the md5 /
symbol counts of the reference's plain-text BWT (`ropebwt2 -L...`, main.c:308-313,323), which is
the canonical, batch- and thread-invariant parity artefact.  Tiny cases also
keep the text itself."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))] + [p for p in sys.path if os.path.abspath(p or '.') != os.path.dirname(os.path.abspath(__file__))]
from oracle.oracle import ref_cli  # noqa: E402
from ropebwt2_b200.synth import from_spec, reads_to_lines  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden.json")


make_reads = from_spec


CASES = [
    # name, generator, flags (reference CLI), strands
    ("cfg1_io", dict(kind="U", n=10000, L=100, seed=42), "-LR"),
    ("cfg1_rlo", dict(kind="U", n=10000, L=100, seed=42), "-LRs"),
    ("cfg1_rclo", dict(kind="U", n=10000, L=100, seed=42), "-LRr"),
    ("cfg1_both_io", dict(kind="U", n=10000, L=100, seed=42), "-L"),
    ("cfg1_both_rclo", dict(kind="U", n=10000, L=100, seed=42), "-Lr"),
    ("cfg1N_rlo", dict(kind="U", n=10000, L=100, seed=43, n_frac=0.001), "-LRs"),
    ("cfg1N_io", dict(kind="U", n=10000, L=100, seed=43, n_frac=0.001), "-LR"),
    ("genome_rlo", dict(kind="G", n=20000, L=101, seed=3), "-LRs"),
    ("genome_both_io", dict(kind="G", n=20000, L=101, seed=3), "-L"),
    ("varlen_io", dict(kind="V", n=3000, L=60, seed=9, lmin=1), "-LR"),
    ("varlen_rlo", dict(kind="V", n=3000, L=60, seed=9, lmin=1), "-LRs"),
    ("varlen_rclo_both", dict(kind="V", n=3000, L=60, seed=9, lmin=1), "-Lr"),
    ("long_io", dict(kind="U", n=200, L=3000, seed=4), "-LR"),
    ("long_rlo", dict(kind="U", n=200, L=3000, seed=4), "-LRs"),
    ("tiny_rlo", dict(kind="V", n=12, L=9, seed=1, lmin=1), "-LRs"),
    ("tiny_io", dict(kind="V", n=12, L=9, seed=1, lmin=1), "-LR"),
]


def main():
    cases = []
    for name, gen, flags in CASES:
        reads = make_reads(gen)
        text = reads_to_lines(reads)
        out, _ = ref_cli([flags, "-"], text)
        # the same input through the single-thread and small-batch paths must agree
        out2, _ = ref_cli([flags + "P", "-m", "20k", "-"], text)
        assert out == out2, name
        body = out[:-1]
        counts = [body.count(c) for c in b"$ACGTN"]
        case = dict(name=name, gen=gen, flags=flags, md5=hashlib.md5(out).hexdigest(), n_symbols=len(body), counts=counts)
        if len(body) <= 400:
            case["text"] = body.decode()
        cases.append(case)
        print(name, case["md5"], len(body))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(dict(reference="lh3/ropebwt2 r187 (bd8dbd3), built unmodified by oracle/Makefile",
                       generator="oracle/gen_golden.py", cases=cases), f, indent=1)


if __name__ == "__main__":
    main()
