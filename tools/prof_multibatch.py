"""Several back-to-back batches into one growing index (the cfg3 shape at reduced size), per-batch statistics:
python tools/prof_multibatch.py [batches] [reads per batch] [workload]"""
import sys, os, ctypes as C, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ropebwt2_b200 import MRope, load, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
per = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
name = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
w = synth.workload(name, nb * per)
L = load()
L.rb2_host_alloc.restype = C.c_void_p
cap = per * (w["L"] + 1)
hptr = L.rb2_host_alloc(cap)
host = np.ctypeslib.as_array((C.c_uint8 * cap).from_address(hptr))
m = MRope(1)
for k in range(nb):
    bench.fill_host_batch(host, w, k * per, (k + 1) * per, torch.device("cuda", 0))
    m.L.mr_insert_multi(m.h, cap, C.cast(hptr, C.POINTER(C.c_uint8)), 1)
    L.rb2_sync(m.engine_handle)   # (the generator runs on the same GPU: keep it out of the insertion's way)
for k, st in enumerate(m.job_history(nb)):
    print(json.dumps({"batch": k, "ms": round(st["ms_total"] - st["ms_h2d"], 1), "ms_merge": round(st["ms_merge"], 1), "merge_GB/s": round(st["merge_bytes_rw"] / max(st["ms_merge"], 1e-9) / 1e6),
                      "ms_groups": round(st["ms_groups"], 1), "ms_members": round(st["ms_members"], 1), "ms_convert": round(st["ms_convert"], 1), "ms_dir": round(st["ms_directory"], 1)}))
