import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ropebwt2_b200 import MRope
from ropebwt2_b200.synth import encode_batch, uniform_reads
so = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for n in [int(x) for x in sys.argv[2:]] or [50000]:
    rd = uniform_reads(n, 101, 1)
    buf = encode_batch(rd)
    for rep in range(2):
        m = MRope(so)
        t = time.time(); m.insert_multi(buf); dt = time.time() - t
        st = m.stats()
        print("%d x 101 so=%d rep%d: %.3fs wall, %.3f Gbp/s" % (n, so, rep, dt, n * 101 / dt / 1e9),
              {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
        m.close()
