"""The counter-based workload generators (ropebwt2_b200/synth.py: hash_reads) exist three times -- numpy
(the definition), torch (bench.py fills the batches on the GPU with it) and C (oracle/gen_reads.c feeds the
reference binary) -- and the md5 parity check at bench size only means something if they agree bit for bit."""
import numpy as np
import pytest

from ropebwt2_b200 import synth


@pytest.mark.parametrize("name", ["U:3000:101:2", "G:3000:101:3", "U:300:2000:4", "cfg5add"])
def test_numpy_torch_c_agree(name):
    import torch
    w = synth.workload(name, 3000 if name == "cfg5add" else 0)
    a = synth.hash_reads(w, 17, w["n"])
    t = synth.hash_reads(w, 17, w["n"], device="cpu").numpy()
    assert a.dtype == np.uint8 and a.min() >= 1 and a.max() <= 4
    assert np.array_equal(a, t)
    lib = synth.c_generator()
    assert lib is not None, "oracle/_build/libgenreads.so missing: make -C oracle oracle"
    c = synth.c_reads(lib, w, 17, w["n"], False, threads=3)
    assert np.array_equal(c[:, :-1], a) and not c[:, -1].any()
    txt = synth.c_reads(lib, w, 17, w["n"], True)
    assert np.array_equal(txt[:, :-1], synth.NT6[a]) and (txt[:, -1] == 10).all()
    # chunking does not matter: a read is a pure function of (seed, index)
    assert np.array_equal(synth.hash_reads(w, 100, 200), a[83:183])


def test_batch_buffers():
    import torch
    w = synth.workload("G:500:60:9")
    buf = np.empty(500 * 61, dtype=np.uint8)
    synth.fill_batch_np(buf, w, 0, 500, chunk=128)
    t = torch.empty(500 * 61, dtype=torch.uint8)
    synth.fill_batch_torch(t, w, 0, 500, chunk=77)
    assert np.array_equal(buf, t.numpy())
    assert np.array_equal(buf, synth.encode_batch(synth.hash_reads(w, 0, 500)))


def test_genome_reads_overlap():
    """kind G really is reads of one genome: at 30x coverage many 20-mers are shared between reads."""
    w = synth.workload("G:4000:101:5")
    r = synth.hash_reads(w, 0, 4000)
    k = {bytes(x[:20]) for x in r} | {bytes((5 - x[::-1])[:20]) for x in r}
    assert sum(bytes(x[i:i + 20]) in k for x in r[:200] for i in range(1, 60)) > 1000
    u = synth.hash_reads(synth.workload("U:4000:101:5"), 0, 4000)
    ku = {bytes(x[:20]) for x in u}
    assert sum(bytes(x[i:i + 20]) in ku for x in u[:200] for i in range(1, 60)) == 0
