"""Timing of the non-headline BASELINE shapes at reduced scale (developer script)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ropebwt2_b200 import MRope
from ropebwt2_b200.synth import encode_batch, uniform_reads, genome_reads
def show(tag, m, bp, dt):
    st = m.stats()
    print(tag, "%.3fs %.3f Gbp/s" % (dt, bp / dt / 1e9), {k: round(v, 1) for k, v in st.items() if k.startswith("ms_")}, "blocks", st["pool_blocks"], flush=True)
# cfg4 shape: long reads, input order
n, L = 20000, 10000
buf = encode_batch(uniform_reads(n, L, 4))
m = MRope(0); m.insert_multi(encode_batch(uniform_reads(100, 100, 1))); m.reset_stats()
t = time.time(); m.insert_multi(buf); show("long reads IO %d x %d" % (n, L), m, n * L, time.time() - t); m.close()
# cfg5 shape: incremental RLO insert into an existing index (non-empty intervals -> rank pre-pass)
base = genome_reads(4_000_000, 101, 3); add = genome_reads(1_000_000, 101, 5)
m = MRope(1); m.insert_multi(encode_batch(base)); m.reset_stats()
t = time.time(); m.insert_multi(encode_batch(add)); show("incremental RLO 1M into 4M (genome reads)", m, add.size, time.time() - t); m.close()
# genome-like data single batch, both strands RCLO
m = MRope(2); b2 = encode_batch(base, True, True); m.reset_stats()
t = time.time(); m.insert_multi(b2); show("genome 4M x 101 both strands RCLO", m, base.size * 2, time.time() - t); m.close()
