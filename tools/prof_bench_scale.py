"""One build at the bench workload size (default 100 M x 101 bp, RLO) for an ncu capture of one late
k_merge_fast launch: python tools/prof_bench_scale.py [reads]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from ropebwt2_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nbytes = n * 102
host = np.empty(nbytes, dtype=np.uint8)
bench.fill_batch(host, n, 101, 2)
e = Engine(0, 1)
d = e.dev_alloc(nbytes); e.dev_upload(d, host)
e.insert_multi_dev(d, nbytes)
st = e.stats()
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
