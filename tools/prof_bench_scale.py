"""One build at the bench workload size (default cfg2: 100 M x 101 bp, RLO) for an ncu capture of one late
k_flat_merge launch: python tools/prof_bench_scale.py [reads]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ropebwt2_b200 import Engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
w = synth.workload("cfg2", n)
nbytes = n * 102
host = np.empty(nbytes, dtype=np.uint8)
bench.fill_host_batch(host, w, 0, n, torch.device("cuda", 0))
e = Engine(0, 1)
d = e.dev_alloc(nbytes); e.dev_upload(d, host)
e.insert_multi_dev(d, nbytes)
st = e.stats()
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
