"""Developer script: dense-regime (RB2_FLAT=1) builds of small inputs against the oracle."""
import os, sys
os.environ.setdefault("RB2_FLAT", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as orc
from ropebwt2_b200 import MRope, load
from ropebwt2_b200.synth import encode_batch, uniform_reads, varlen_reads

def run(name, so, bufs):
    o, m = orc.Oracle(so), MRope(so)
    for i, b in enumerate(bufs):
        o.insert_multi(b); m.insert_multi(b)
        got = orc.decode_index(load(), m.h, m.total())[0]
        want = o.text()
        ok = got.shape == want.shape and np.array_equal(got, want)
        msg = ""
        if not ok:
            if got.shape != want.shape: msg = f"sizes {got.shape} vs {want.shape}"
            else:
                d = np.nonzero(got != want)[0]
                msg = f"{len(d)} diffs, first at {d[0]}: got {got[d[0]:d[0]+12]} want {want[d[0]:d[0]+12]}"
        print(f"{name} so={so} batch {i}: total {m.total()} {'OK' if ok else 'MISMATCH ' + msg}", flush=True)
        if not ok: break
    m.close()

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("tiny", "all"):
    run("tiny", 0, [encode_batch(uniform_reads(5, 8, 1))])
    run("tiny", 1, [encode_batch(uniform_reads(5, 8, 1))])
if which in ("small", "all"):
    for so in (0, 1, 2):
        run("small", so, [encode_batch(uniform_reads(3000, 40, 7, n_frac=0.01))])
if which in ("multi", "all"):
    for so in (0, 1, 2):
        rd = uniform_reads(9000, 30, 3)
        run("multi", so, [encode_batch(rd[:3000]), encode_batch(rd[3000:5000]), encode_batch(rd[5000:])])
if which in ("long", "all"):
    rd = np.tile(np.array([[1, 2, 2, 4]], dtype=np.uint8), (600000, 1))
    run("longruns", 1, [encode_batch(rd), encode_batch(uniform_reads(100, 6, 1))])
if which in ("var", "all"):
    run("var", 1, [encode_batch(varlen_reads(700, 60, 5), True, True), encode_batch(varlen_reads(300, 20, 6), True, True)])
