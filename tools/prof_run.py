"""One build of n x 101 bp reads for profiling under ncu: python tools/prof_run.py <n> [so]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ropebwt2_b200 import MRope
from ropebwt2_b200.synth import encode_batch, uniform_reads
n = int(sys.argv[1]); so = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = MRope(so)
m.insert_multi(encode_batch(uniform_reads(n, 101, 1)))
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in m.stats().items()})
