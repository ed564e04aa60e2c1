"""Developer script (not a test): progressively larger parity checks against the oracle with
verbose mismatch output.  Run on a GPU box: python tools/gpu_debug.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle, decode_index
from ropebwt2_b200 import MRope, load
from ropebwt2_b200.synth import encode_batch, uniform_reads, genome_reads

L = load()
S = "$ACGTN"

def show(t): return "".join(S[c] for c in t[:200])

def check(name, so, batches):
    o = Oracle(so); m = MRope(so)
    for buf in batches:
        o.insert_multi(buf); m.insert_multi(buf)
    tot = o.total()
    ok_cnt = np.array_equal(o.counts(), m.counts())
    try:
        tg = decode_index(L, m.h, tot)[0]
        to = o.text()
        ok = np.array_equal(tg, to)
    except AssertionError as e:
        ok = False; tg = None; to = o.text(); print("  decode failed:", e)
    print(("OK  " if ok and ok_cnt else "FAIL"), name, "so", so, "sym", tot, "counts_ok", ok_cnt, flush=True)
    if not (ok and ok_cnt):
        if tg is not None:
            d = np.nonzero(tg != to)[0]
            print("  first diff at", d[:5], "of", tot)
            if tot <= 200: print("  gpu", show(tg)); print("  ora", show(to))
        print("  oracle counts\n", o.counts(), "\n  gpu counts\n", m.counts())
        return False
    return True

rng = np.random.default_rng(5)
allok = True
# tiny
for it in range(60):
    so = it % 3
    n = int(rng.integers(1, 10))
    strs = [rng.integers(1, 6 if it % 4 == 0 else 5, size=int(rng.integers(0, 8))).astype(np.uint8) for _ in range(n)]
    if it % 5 == 0: strs += [strs[0].copy()]
    nb = 1 + it % 3
    cuts = sorted(rng.integers(0, len(strs) + 1, size=nb - 1).tolist())
    parts = [strs[a:b] for a, b in zip([0] + cuts, cuts + [len(strs)])]
    bufs = [encode_batch(p, True, it % 2 == 1) for p in parts if p]
    if not check(f"tiny{it} n={len(strs)} nb={nb}", so, bufs):
        print("   strings:", [s.tolist() for s in strs], "cuts", cuts)
        allok = False
        break
if allok:
    for so in (0, 1, 2):
        for (n, l, nb) in ((200, 30, 1), (2000, 50, 2), (10000, 100, 1), (10000, 100, 3), (300, 1500, 2)):
            rd = uniform_reads(n, l, 11 + so, n_frac=0.01)
            step = (n + nb - 1) // nb
            bufs = [encode_batch(rd[a:a + step], True, False) for a in range(0, n, step)]
            t = time.time()
            allok &= check(f"U n={n} l={l} nb={nb}", so, bufs)
            print("     %.2fs" % (time.time() - t))
        rd = genome_reads(20000, 101, 3, coverage=30)
        allok &= check("G 20000x101 3 batches", so, [encode_batch(rd[a:a + 7000], True, True) for a in range(0, 20000, 7000)])
print("ALL OK" if allok else "SOME FAILED")
m = MRope(1); rd = uniform_reads(200000, 101, 1)
buf = encode_batch(rd)
t = time.time(); m.insert_multi(buf); dt = time.time() - t
st = m.stats()
print("200k x 101 RLO: %.3fs wall" % dt, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
