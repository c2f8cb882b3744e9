"""CPU test: the index builder reproduces the reference's shipped index byte for byte (pins the .bwt/.sa/.pac/.ann/
.amb encodings the alignment core reads), and the .map image round-trips through bwa_idx2mem / bwa_mem2idx."""
import os
import numpy as np
from mpibwa_b200 import index_build, simulate


def test_rebuild_hg19_small(examples, tmp_path):
    """one bucket range, and many small ones (the way a human-sized reference is built: memory bounded by the range size)"""
    fa = examples["idx"]
    for batch in (None, 300_000):
        prefix = str(tmp_path / ("rebuilt%s.fa" % batch))
        index_build.build_index(fa, prefix, device="cpu", batch=batch)
        for ext in (".pac", ".ann", ".amb", ".bwt", ".sa"):
            assert open(prefix + ext, "rb").read() == open(fa + ext, "rb").read(), (ext, batch)


def test_suffix_array_small_texts():
    rng = np.random.default_rng(3)
    for n, alphabet in ((1, 4), (2, 4), (40, 4), (200, 1), (300, 2), (1000, 4)):
        t = rng.integers(0, alphabet, size=n).astype(np.uint8)
        if n == 300:
            t[100:200] = t[0:100]          # long exact repeat -> several refinement rounds
        s = bytes(t + 1)
        want = sorted(range(n), key=lambda i: s[i:])
        for batch in (None, 7):
            assert index_build.suffix_array(t, device="cpu", batch=batch).tolist() == want, (n, batch)


def test_simulator_is_seeded():
    names, lengths, codes = simulate.make_reference(50000, 3, seed=7)
    a = simulate.simulate_pairs(codes, lengths, 300, seed=9)
    b = simulate.simulate_pairs(codes, lengths, 300, seed=9)
    assert a == b and a[0].count(b"\n") == 1200 and int(lengths.sum()) == 50000
