"""CPU tests: pin oracle/liboracle.so (our C restatement) against the UNMODIFIED reference compiled into
oracle/_ref/libbwa_ref.so, function by function, and against the committed golden vectors (tests/golden/*.json,
produced from the reference by tests/golden/make_golden.py)."""
import ctypes as C
import json
import os
import numpy as np
import pytest
from conftest import ROOT, have_ref
import fuzzgen
import oracle_lib as OL

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (reference sources absent)")


@pytest.fixture(scope="module")
def orc(oracle_built):
    return OL.Oracle()


@pytest.fixture(scope="module")
def ref(oracle_built):
    return OL.Reference()


@needs_ref
def test_extend_matches_reference(orc, ref):
    for c in fuzzgen.extend_cases(11, 4000):
        a, b, od, ed, oi, ei = c["params"]
        mat = OL.default_mat(a, b)
        args = (c["q"], c["t"], mat, od, ed, oi, ei, c["w"], c["end_bonus"], c["zdrop"], c["h0"])
        assert orc.extend(*args)[0] == ref.extend(*args), c


@needs_ref
@pytest.mark.parametrize("sixteen", [False, True])
def test_align_matches_reference(orc, ref, sixteen):
    for c in fuzzgen.align_cases(21 + sixteen, 1500, sixteen):
        a, b, od, ed, oi, ei = c["params"]
        mat = OL.default_mat(a, b)
        args = (c["q"], c["t"], mat, od, ed, oi, ei, c["xtra"])
        assert orc.align(*args)[0] == ref.align(*args), c


def test_global_matches_reference(orc, ref):
    """orc_ksw_global2 (full-matrix restatement) against the reference's ksw_global2: score and CIGAR"""
    for c in fuzzgen.global_cases(61, 1500):
        a, b, od, ed, oi, ei = c["params"]
        mat = OL.default_mat(a, b)
        assert orc.global2(c["q"], c["t"], mat, od, ed, oi, ei, c["w"])[0] == ref.global2(c["q"], c["t"], mat, od, ed, oi, ei, c["w"]), (len(c["q"]), len(c["t"]), c["w"])


def test_golden_vectors(orc):
    """golden input/output vectors recorded from the reference (travel to machines without the reference)"""
    with open(os.path.join(ROOT, "tests", "golden", "ksw_vectors.json")) as fh:
        g = json.load(fh)
    for v in g["extend"]:
        mat = OL.default_mat(v["a"], v["b"])
        got = orc.extend(bytes(v["q"]), bytes(v["t"]), mat, v["o_del"], v["e_del"], v["o_ins"], v["e_ins"], v["w"], v["end_bonus"], v["zdrop"], v["h0"])[0]
        assert list(got) == v["out"]
    for v in g["align"]:
        mat = OL.default_mat(v["a"], v["b"])
        got = orc.align(bytes(v["q"]), bytes(v["t"]), mat, v["o_del"], v["e_del"], v["o_ins"], v["e_ins"], v["xtra"])[0]
        assert list(got) == v["out"]


@needs_ref
def test_fm_index_matches_reference(orc, ref, examples):
    idxf = OL.IndexFiles(examples["idx"])
    ridx = ref.lib.bwa_idx_load(examples["idx"].encode(), 7)
    bwt = ridx.contents.bwt
    rng = np.random.default_rng(5)
    # reads sampled from the reference text (both strands) with a few errors so that SMEMs are non-trivial
    codes = np.zeros(idxf.l_pac, dtype=np.uint8)
    pac = idxf.pac
    pos = np.arange(idxf.l_pac)
    codes = (pac[pos >> 2] >> ((~pos & 3) << 1)) & 3
    n_smem = 0
    for it in range(150):
        L = int(rng.integers(30, 200))
        s = int(rng.integers(0, idxf.l_pac - L))
        q = fuzzgen.mutate(rng, codes[s:s + L], 0.03, 0.005, 0.01)
        if rng.random() < 0.5:
            q = np.where(q < 4, 3 - q, 4)[::-1].astype(np.uint8)
        if rng.random() < 0.1:
            q = rng.integers(0, 4, size=L).astype(np.uint8)
        qb = bytes(q)
        # bwt_smem1 at every start position the reference's pass 1 would visit, plus pass-2 style calls
        x = 0
        while x < len(q):
            if q[x] > 3:
                x += 1
                continue
            r_ret, r_mem = ref.smem1(bwt, q, x, 1)
            mem = (OL.orc_intv_t * (len(q) + 1))()
            n = C.c_int()
            o_ret = orc.lib.orc_smem1(C.byref(idxf.fm), len(q), qb, x, 1, mem, C.byref(n))
            o_mem = [(mem[i].x0, mem[i].x1, mem[i].x2, mem[i].info) for i in range(n.value)]
            assert (o_ret, o_mem) == (r_ret, r_mem)
            for (x0, x1, x2, info) in r_mem[:2]:
                mid = ((info >> 32) + (info & 0xffffffff)) >> 1
                r2 = ref.smem1(bwt, q, mid, x2 + 1)
                n2 = C.c_int()
                o2 = orc.lib.orc_smem1(C.byref(idxf.fm), len(q), qb, mid, x2 + 1, mem, C.byref(n2))
                assert (o2, [(mem[i].x0, mem[i].x1, mem[i].x2, mem[i].info) for i in range(n2.value)]) == r2
            n_smem += len(r_mem)
            x = r_ret
        # bwt_seed_strategy1
        x = 0
        while x < len(q):
            if q[x] > 3:
                x += 1
                continue
            rm = OL.ref_bwtintv_t()
            om = OL.orc_intv_t()
            rr = ref.lib.bwt_seed_strategy1(bwt, len(q), qb, x, 19, 20, C.byref(rm))
            orr = orc.lib.orc_seed_strategy1(C.byref(idxf.fm), len(q), qb, x, 19, 20, C.byref(om))
            assert rr == orr and (rm.x[2] == om.x2) and (rm.x[2] == 0 or (rm.x[0], rm.x[1], rm.info) == (om.x0, om.x1, om.info))
            x = rr
    assert n_smem > 100
    # suffix array look-ups, including the row of the sentinel and sampled rows
    ks = [1, idxf.primary, idxf.seq_len, 32, 64] + [int(v) for v in rng.integers(1, idxf.seq_len + 1, size=3000)]
    for k in ks:
        assert orc.lib.orc_sa(C.byref(idxf.fm), k) == ref.lib.bwt_sa(bwt, k)
    # reference windows on both strands and across the strand boundary
    for it in range(300):
        b = int(rng.integers(0, 2 * idxf.l_pac))
        e = min(2 * idxf.l_pac, b + int(rng.integers(0, 400)))
        if it % 10 == 0:
            b, e = idxf.l_pac - 50, idxf.l_pac + 50
        ln = C.c_int64()
        p = ref.lib.bns_get_seq(idxf.l_pac, ridx.contents.pac, b, e, C.byref(ln))
        want = bytes(p[:ln.value]) if ln.value else b""
        ref.libc.free(p)
        buf = np.zeros(e - b + 1, dtype=np.uint8)
        n = orc.lib.orc_get_seq(idxf.l_pac, idxf.pac.ctypes.data, b, e, buf.ctypes.data)
        assert bytes(buf[:n]) == want
