"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI of libmpibwa_b200.so and is
compared bit-exactly (integer/byte work: no tolerance) with
  * oracle/liboracle.so, our C restatement pinned against the reference (kernel level),
  * the golden vectors and SAM digests recorded from the reference (tests/golden),
  * oracle/_ref/ref_driver, the compiled reference itself, when it travelled to the box (end to end)."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import numpy as np
import pytest
from conftest import ROOT, have_ref, have_cuda
import fuzzgen
import oracle_lib as OL
import mpibwa_b200 as M

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="no CUDA device")]
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def orc(oracle_built):
    return OL.Oracle()


@pytest.fixture(scope="module")
def aligner(examples):
    a = M.Aligner(examples["idx"], device=0, n_threads=8, verbose=1)
    yield a


def _run_extend_batch(lib, cases):
    """group by (params, zdrop) because those are per-call arguments of the batch entry point"""
    out = [None] * len(cases)
    groups = {}
    for i, c in enumerate(cases):
        groups.setdefault((c["params"], c["zdrop"]), []).append(i)
    for (params, zdrop), idxs in groups.items():
        a, b, od, ed, oi, ei = params
        jobs = (M.b200_extend_job_t * len(idxs))()
        qs, ts, qo, to = [], [], 0, 0
        for k, i in enumerate(idxs):
            c = cases[i]
            jobs[k].qlen, jobs[k].tlen, jobs[k].q_off, jobs[k].t_off = len(c["q"]), len(c["t"]), qo, to
            jobs[k].h0, jobs[k].w, jobs[k].end_bonus = c["h0"], c["w"], c["end_bonus"]
            qs.append(bytes(c["q"])); ts.append(bytes(c["t"]))
            qo += len(c["q"]); to += len(c["t"])
        qb, tb = b"".join(qs), b"".join(ts)
        lib.b200_ksw_extend2_batch(len(idxs), jobs, qb, len(qb), tb, len(tb), OL.default_mat(a, b), od, ed, oi, ei, zdrop)
        for k, i in enumerate(idxs):
            j = jobs[k]
            out[i] = (j.score, j.qle, j.tle, j.gtle, j.gscore, j.max_off)
    return out


def _run_align_batch(lib, cases):
    out = [None] * len(cases)
    groups = {}
    for i, c in enumerate(cases):
        groups.setdefault(c["params"], []).append(i)
    for params, idxs in groups.items():
        a, b, od, ed, oi, ei = params
        jobs = (M.b200_align_job_t * len(idxs))()
        qs, ts, qo, to = [], [], 0, 0
        for k, i in enumerate(idxs):
            c = cases[i]
            jobs[k].qlen, jobs[k].tlen, jobs[k].q_off, jobs[k].t_off, jobs[k].xtra = len(c["q"]), len(c["t"]), qo, to, c["xtra"]
            qs.append(bytes(c["q"])); ts.append(bytes(c["t"]))
            qo += len(c["q"]); to += len(c["t"])
        qb, tb = b"".join(qs), b"".join(ts)
        lib.b200_ksw_align2_batch(len(idxs), jobs, qb, len(qb), tb, len(tb), OL.default_mat(a, b), od, ed, oi, ei)
        for k, i in enumerate(idxs):
            r = jobs[k].r
            out[i] = (r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb)
    return out


def _aux_stats(lib):
    st = M.b200_stats_t()
    lib.b200_get_aux_stats(C.byref(st))
    return st.as_dict()


@pytest.mark.parametrize("kernel", ["auto", "lane", "warp", "big"])
def test_extend_batch_vs_oracle(aligner, orc, kernel):
    """ksw_extend2 fuzz through the C ABI; every DP kernel of the extension stage is forced in turn (one job per lane with
    packed 16-bit rows in shared memory, one warp per job with the max-plus scan, general int32 rows in global memory)"""
    cases = fuzzgen.extend_cases(31, 12000) + fuzzgen.extend_cases(32, 2000, max_q=250, max_t=1100) + \
        fuzzgen.extend_cases(33, 300, max_q=900, max_t=1500)
    os.environ.pop("B200_EXT_KERNEL", None)
    if kernel != "auto":
        os.environ["B200_EXT_KERNEL"] = kernel
    st0 = _aux_stats(aligner.lib)
    try:
        got = _run_extend_batch(aligner.lib, cases)
    finally:
        os.environ.pop("B200_EXT_KERNEL", None)
    st1 = _aux_stats(aligner.lib)
    cells = 0
    for c, g in zip(cases, got):
        a, b, od, ed, oi, ei = c["params"]
        want, n_cells = orc.extend(c["q"], c["t"], OL.default_mat(a, b), od, ed, oi, ei, c["w"], c["end_bonus"], c["zdrop"], c["h0"])
        assert g == want, c
        cells += n_cells
    # the device's cell counter (the numerator of the ksw_extend2 roofline) counts what the reference executes: sum over rows of end - beg
    assert st1["extend_cells"] - st0["extend_cells"] == cells


@pytest.mark.parametrize("sixteen", [False, True])
def test_align_batch_vs_oracle(aligner, orc, sixteen):
    cases = fuzzgen.align_cases(41 + sixteen, 6000, sixteen)
    st0 = _aux_stats(aligner.lib)
    got = _run_align_batch(aligner.lib, cases)
    st1 = _aux_stats(aligner.lib)
    cells = 0
    for c, g in zip(cases, got):
        a, b, od, ed, oi, ei = c["params"]
        want, n_cells = orc.align(c["q"], c["t"], OL.default_mat(a, b), od, ed, oi, ei, c["xtra"])
        assert g == want, c
        cells += n_cells
    assert st1["sw_cells"] - st0["sw_cells"] == cells          # padded query length x rows executed, both passes


@pytest.mark.parametrize("squeeze", [False, True])
def test_global_batch_vs_oracle(aligner, orc, squeeze, monkeypatch):
    """ksw_global2 fuzz through the kernels of the CIGAR stage (k_global_lanes: band-wide circular row window in shared memory,
    classes by window width; k_global_jobs: general path): score, CIGAR and cell count against the oracle's orc_ksw_global2 for
    bands 1..200, tlen != qlen and gappy pairs with dozens of operations; squeeze: windows too small for the band, every job takes
    the rerun pass"""
    if squeeze:
        monkeypatch.setenv("B200_GLOBAL_SQUEEZE", "1")
    cases = fuzzgen.global_cases(71, 6000) + fuzzgen.global_cases(72, 300, max_q=600)
    for c in cases:
        c["t"] = c["t"] & 3
    groups = {}
    for i, c in enumerate(cases):
        groups.setdefault(c["params"], []).append(i)
    st0 = _aux_stats(aligner.lib)
    got, cells, max_ops = {}, 0, 0
    for params, idxs in groups.items():
        a, b, od, ed, oi, ei = params
        jobs = (M.b200_global_job_t * len(idxs))()
        qs, ts, qo, to = [], [], 0, 0
        for k, i in enumerate(idxs):
            c = cases[i]
            jobs[k].qlen, jobs[k].tlen, jobs[k].q_off, jobs[k].t_off, jobs[k].w = len(c["q"]), len(c["t"]), qo, to, c["w"]
            qs.append(bytes(c["q"])); ts.append(bytes(c["t"]))
            qo += len(c["q"]); to += len(c["t"])
        qb, tb = b"".join(qs), b"".join(ts)
        cig = C.POINTER(C.c_uint32)()
        aligner.lib.b200_ksw_global2_batch(len(idxs), jobs, qb, len(qb), tb, len(tb), OL.default_mat(a, b), od, ed, oi, ei, C.byref(cig))
        for k, i in enumerate(idxs):
            got[i] = (jobs[k].score, [cig[jobs[k].cigar_off + x] for x in range(jobs[k].n_cigar)])
        aligner.lib.b200_free(cig)
    st1 = _aux_stats(aligner.lib)
    for i, c in enumerate(cases):
        a, b, od, ed, oi, ei = c["params"]
        want, n_cells = orc.global2(c["q"], c["t"], OL.default_mat(a, b), od, ed, oi, ei, c["w"])
        assert got[i] == want, (i, len(c["q"]), len(c["t"]), c["w"])
        cells += n_cells
        max_ops = max(max_ops, len(want[1]))
    assert max_ops > 28                                   # (round 1's result record held 28 operations)
    if not squeeze:
        assert st1["global_cells"] - st0["global_cells"] == cells


def test_golden_vectors_through_cabi(aligner):
    g = json.load(open(os.path.join(GOLD, "ksw_vectors.json")))
    ext = [dict(q=np.array(v["q"], np.uint8), t=np.array(v["t"], np.uint8), params=(v["a"], v["b"], v["o_del"], v["e_del"], v["o_ins"], v["e_ins"]),
                w=v["w"], end_bonus=v["end_bonus"], zdrop=v["zdrop"], h0=v["h0"]) for v in g["extend"]]
    assert [list(x) for x in _run_extend_batch(aligner.lib, ext)] == [v["out"] for v in g["extend"]]
    aln = [dict(q=np.array(v["q"], np.uint8), t=np.array(v["t"], np.uint8), params=(v["a"], v["b"], v["o_del"], v["e_del"], v["o_ins"], v["e_ins"]),
                xtra=v["xtra"]) for v in g["align"]]
    assert [list(x) for x in _run_align_batch(aligner.lib, aln)] == [v["out"] for v in g["align"]]


def test_single_job_wrappers(aligner, orc):
    """the reference's own call surface (ksw_extend2 / ksw_align2 / bwt_sa / bwt_extend) as batches of one"""
    lib = aligner.lib
    idxf = OL.IndexFiles(aligner_idx_prefix(aligner))
    for c in fuzzgen.extend_cases(51, 20):
        a, b, od, ed, oi, ei = c["params"]
        o = [C.c_int() for _ in range(5)]
        sc = lib.ksw_extend2(len(c["q"]), bytes(c["q"]), len(c["t"]), bytes(c["t"]), 5, OL.default_mat(a, b), od, ed, oi, ei, c["w"],
                             c["end_bonus"], c["zdrop"], c["h0"], *[C.byref(x) for x in o])
        want = orc.extend(c["q"], c["t"], OL.default_mat(a, b), od, ed, oi, ei, c["w"], c["end_bonus"], c["zdrop"], c["h0"])[0]
        assert (sc, *[x.value for x in o]) == want
    for k in (1, 33, idxf.primary, idxf.seq_len):
        assert lib.bwt_sa(aligner.idx.contents.bwt, k) == orc.lib.orc_sa(C.byref(idxf.fm), k)
    ik = M.bwtintv_t((C.c_uint64 * 3)(idxf.L2[1] + 1, idxf.L2[2] + 1, idxf.L2[2] - idxf.L2[1]), 0)
    ok = (M.bwtintv_t * 4)()
    lib.bwt_extend(aligner.idx.contents.bwt, C.byref(ik), ok, 1)
    oik = OL.orc_intv_t(ik.x[0], ik.x[1], ik.x[2], 0)
    ook = (OL.orc_intv_t * 4)()
    orc.lib.orc_extend(C.byref(idxf.fm), C.byref(oik), ook, 1)
    assert [(ok[i].x[0], ok[i].x[1], ok[i].x[2]) for i in range(4)] == [(ook[i].x0, ook[i].x1, ook[i].x2) for i in range(4)]
    # bwt_smem1: all SMEMs through one query position (served by the device, reference src/bwt.c:353)
    reads = _reads_as_codes(_PREFIX["R1"], 12)
    for q in reads:
        for x in (0, len(q) // 3, len(q) - 1):
            for min_intv in (1, 3):
                mem = OL.ref_bwtintv_v()
                ret = lib.bwt_smem1(aligner.idx.contents.bwt, len(q), bytes(q), x, min_intv, C.byref(mem), None)
                out = (OL.orc_intv_t * (len(q) + 1))()
                n = C.c_int()
                want_ret = orc.lib.orc_smem1(C.byref(idxf.fm), len(q), bytes(q), x, min_intv, out, C.byref(n))
                assert ret == want_ret and mem.n == n.value
                assert [(mem.a[i].x[0], mem.a[i].x[1], mem.a[i].x[2], mem.a[i].info) for i in range(mem.n)] == \
                    [(out[i].x0, out[i].x1, out[i].x2, out[i].info) for i in range(n.value)]
                lib.b200_free(mem.a)
    # ksw_global2: score and CIGAR of a banded global alignment (served by the device, reference src/ksw.c:504)
    if have_ref():
        ref = OL.Reference()
        ref.lib.ksw_global2.argtypes = lib.ksw_global2.argtypes
        for c in fuzzgen.global_cases(77, 60):            # (bands that reach the last cell: |tlen - qlen| <= w, as every caller guarantees)
            q, t, w = c["q"], c["t"] & 3, c["w"]
            outs = []
            for L in (lib, ref.lib):
                n_cigar, cigar = C.c_int(), C.POINTER(C.c_uint32)()
                sc = L.ksw_global2(len(q), bytes(q), len(t), bytes(t), 5, OL.default_mat(1, 4), 6, 1, 5, 2, w, C.byref(n_cigar), C.byref(cigar))
                outs.append((sc, [cigar[i] for i in range(n_cigar.value)]))
            assert outs[0] == outs[1], (len(q), len(t), w)


_PREFIX = {}


def aligner_idx_prefix(aligner):
    return _PREFIX["idx"]


@pytest.fixture(autouse=True)
def _remember_prefix(examples):
    _PREFIX["idx"] = examples["idx"]
    _PREFIX["R1"] = examples["R1_10K"]


def _reads_as_codes(path, n):
    nt4 = np.full(256, 4, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        nt4[ch] = i; nt4[ch + 32] = i
    out = []
    with open(path, "rb") as fh:
        for k, line in enumerate(fh):
            if k % 4 == 1:
                out.append(nt4[np.frombuffer(line.rstrip(b"\n"), np.uint8)])
                if len(out) == n:
                    break
    return out


_occ_blocks = {}


@pytest.mark.parametrize("kernel", ["sweeps", "lanes", "sweeps_overflow", "sweeps_no_tables", "sweeps_tables5"])
def test_seeding_and_sa_vs_oracle(aligner, orc, examples, kernel, monkeypatch):
    # the homogeneous sweep kernels (default: short patterns served by the k-mer interval tables), the general per-lane state
    # machine (extensions only), sweeps with strips so small that most reads overflow and are redone by the general kernel, and the
    # sweeps without tables / with tables of only five bases (an index uploaded anew with B200_KMER_MAX)
    if kernel == "lanes":
        monkeypatch.setenv("B200_SEED_KERNEL", "lanes")
    if kernel == "sweeps_overflow":
        monkeypatch.setenv("B200_SEED_STRIP", "40")
    if kernel in ("sweeps_no_tables", "sweeps_tables5"):
        monkeypatch.setenv("B200_KMER_MAX", "0" if kernel == "sweeps_no_tables" else "5")
        if kernel == "sweeps_no_tables":
            monkeypatch.setenv("B200_SA_FULL", "0")        # ... and bwt_sa by the walk to a sampled row instead of the expanded array
        else:
            monkeypatch.setenv("B200_BLOOM", "0")          # ... and no Bloom filter in front of the backward chains
        aligner = M.Aligner(examples["idx"], device=0, n_threads=8, verbose=1)
    blocks0 = _aux_stats(aligner.lib)["fm_occ_blocks"]
    idxf = OL.IndexFiles(examples["idx"])
    rng = np.random.default_rng(9)
    reads = _reads_as_codes(examples["R1_10K"], 400)
    pos = np.arange(idxf.l_pac)
    text = (idxf.pac[pos >> 2] >> ((~pos & 3) << 1)) & 3
    for _ in range(400):
        L = int(rng.integers(10, 260))
        s = int(rng.integers(0, idxf.l_pac - L))
        q = fuzzgen.mutate(rng, text[s:s + L], 0.02, 0.004, 0.005)
        reads.append(q if rng.random() < 0.5 else np.where(q < 4, 3 - q, 4)[::-1].astype(np.uint8))
    reads.append(np.full(50, 4, np.uint8))                 # all N
    reads.append(np.zeros(5, np.uint8))                    # shorter than a seed
    reads.append(np.zeros(0, np.uint8))                    # empty
    off = np.zeros(len(reads) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    flat = np.concatenate(reads + [np.zeros(8, np.uint8)])
    intv = C.POINTER(M.bwtintv_t)()
    ioff = C.POINTER(C.c_int64)()
    aligner.lib.b200_collect_intv_batch(aligner.opt, len(reads), off.ctypes.data, flat.ctypes.data, C.byref(intv), C.byref(ioff))
    n_total = 0
    for r, q in enumerate(reads):
        got = [(intv[i].x[0], intv[i].x[1], intv[i].x[2], intv[i].info) for i in range(ioff[r], ioff[r + 1])]
        assert got == orc.collect_intv(idxf.fm, q), r
        n_total += len(got)
    assert n_total > 1000
    aligner.lib.b200_free(intv); aligner.lib.b200_free(ioff)
    # occ blocks touched: the general kernel runs the reference's loops (its count is the algorithmic traffic of SURVEY.md 8d); the
    # sweeps look short patterns up and walk the backward entries as independent chains, which must not cost more than the loops
    if kernel != "sweeps_overflow":
        _occ_blocks[kernel] = _aux_stats(aligner.lib)["fm_occ_blocks"] - blocks0
        assert _occ_blocks[kernel] > 100000
        if "lanes" in _occ_blocks and "sweeps" in _occ_blocks:
            assert _occ_blocks["sweeps"] < 0.7 * _occ_blocks["lanes"], _occ_blocks
        if "lanes" in _occ_blocks and "sweeps_no_tables" in _occ_blocks:
            assert _occ_blocks["sweeps_no_tables"] < 1.1 * _occ_blocks["lanes"], _occ_blocks
    ks = np.concatenate([[1, idxf.primary, idxf.seq_len, 32], rng.integers(1, idxf.seq_len + 1, size=20000)]).astype(np.uint64)
    sa = np.zeros(len(ks), np.uint64)
    aligner.lib.b200_bwt_sa_batch(len(ks), ks.ctypes.data, sa.ctypes.data)
    want = [orc.lib.orc_sa(C.byref(idxf.fm), int(k)) for k in ks[:3000]]
    assert sa[:3000].tolist() == want
    # size-independent property: SA values of distinct rows are distinct text positions
    assert len(set(sa.tolist())) == len(set(ks.tolist()))


@pytest.mark.parametrize("dom", [1, 3])
def test_fm_index_beyond_2_32(tmp_path, orc, dom):
    """occ sectors at BWT rows AND symbol counts beyond 2^32 (human-sized references: 2 x 3.1 Gbp rows), on the real device path:
    a synthetic 4.4 G-row index in the reference's file formats (tests/bigfm.py) is loaded through bwa_idx_load, re-blocked into
    occ sectors by the upload kernel like any index, and bwt_extend (ld_occ's 256-bit sector load + occ4_sector's 40-bit counts,
    the primitive of every seeding and SA kernel) is compared with the oracle's orc_extend over the same files, for intervals
    that start, end and straddle rows around 2^32 and hold up to 2^32 + 999 rows.  (The symbol array is not the transform of a
    text, so only single steps are compared: a sweep over it would leave the range of valid rows.)"""
    import bigfm
    prefix = str(tmp_path / "big.fa")
    info = bigfm.write_index(prefix, dom=dom, seed=5 + dom)
    idxf = OL.IndexFiles(prefix)
    a = M.Aligner(prefix, device=0, n_threads=4, verbose=1)
    try:
        lib = a.lib
        rng = np.random.default_rng(13)
        n_bad = 0
        for t in range(400):
            w0, w1 = info["windows"][t % len(info["windows"])]
            k = int(rng.integers(w0 - 300, w1 + 300))
            size = int(rng.choice([1, 2, 50, 4000, 1 << 20, 1 << 31, (1 << 32) + 999]))
            size = min(size, info["seq_len"] - k)
            ik = M.bwtintv_t((C.c_uint64 * 3)(k, int(rng.integers(1, info["seq_len"] - size)), size), 0)
            for is_back in (0, 1):
                ok = (M.bwtintv_t * 4)()
                lib.bwt_extend(a.idx.contents.bwt, C.byref(ik), ok, is_back)
                oik = OL.orc_intv_t(ik.x[0], ik.x[1], ik.x[2], 0)
                ook = (OL.orc_intv_t * 4)()
                orc.lib.orc_extend(C.byref(idxf.fm), C.byref(oik), ook, is_back)
                n_bad += [(ok[i].x[0], ok[i].x[1], ok[i].x[2]) for i in range(4)] != [(ook[i].x0, ook[i].x1, ook[i].x2) for i in range(4)]
        assert n_bad == 0
    finally:
        a.close()


def _sq_header(aligner):
    bns = aligner.idx.contents.bns.contents
    return b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (bns.anns[i].name, bns.anns[i].len) for i in range(bns.n_seqs))


@pytest.mark.parametrize("shape", ["pe", "pe_K", "trim", "se"])
def test_examples_sam_digest(examples, shape):
    """bit-exact SAM vs the digests recorded from the reference on its own example data"""
    want = json.load(open(os.path.join(GOLD, "sam_md5.json")))[shape]
    rd = lambda k: open(examples[k], "rb").read()
    a = M.Aligner(examples["idx"], device=0, n_threads=8, paired=(shape != "se"), verbose=1)
    if shape == "pe":
        sam = _sq_header(a) + a.align(rd("R1_10K"), rd("R2_10K"), K=10000000 * 8)
    elif shape == "pe_K":
        sam = a.align(rd("R1_10K"), rd("R2_10K"), K=500000)
    elif shape == "trim":
        sam = a.align(rd("R1_10K_TRIM"), rd("R2_10K_TRIM"), K=700000, trimmed=True)
    else:
        sam = a.align(rd("R1_10K"), None, K=300000)
    assert sam.count(b"\n") == want["lines"]
    assert hashlib.md5(sam).hexdigest() == want["md5"]
    st = a.stats()
    assert st["n_launches"] > 0 and st["extend_cells"] > 0


SYN = {  # name: (ref bp, contigs, simulate kwargs, driver args)
    "cfg2_like": (3_000_000, 4, dict(n_pairs=30000), ["-K", "4000000"]),
    "cfg4_like": (2_000_000, 3, dict(n_pairs=8000, read_len=250, sub=0.03, indel=0.01, max_indel=50, trim_to=(100, 250),
                                     unmappable_frac=0.2, seed=5), ["-T", "-K", "900000"]),
    "se100": (2_000_000, 5, dict(n_pairs=20000, read_len=100, seed=8), ["-K", "500000"]),
}


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("name", list(SYN))
def test_synthetic_vs_compiled_reference(tmp_path, name):
    """fresh synthetic reference + wgsim-style reads: SAM of the B200 core == SAM of the compiled reference"""
    from mpibwa_b200 import simulate, index_build
    bp, nctg, kw, args = SYN[name]
    names, lengths, codes = simulate.make_reference(bp, nctg, seed=17)
    prefix = str(tmp_path / "ref.fa")
    index_build.write_fasta(prefix, simulate.codes_to_fasta_contigs(names, lengths, codes, n_runs=2))
    index_build.build_index(prefix)
    r1, r2 = simulate.simulate_pairs(codes, lengths, **kw)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    open(f1, "wb").write(r1); open(f2, "wb").write(r2)
    fq = [f1] if name == "se100" else [f1, f2]
    want = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "-t", "16"] + args + [prefix] + fq,
                          capture_output=True, check=True).stdout
    a = M.Aligner(prefix, device=0, n_threads=16, paired=(name != "se100"), verbose=1)
    K = int(args[args.index("-K") + 1])
    got = a.align(r1, None if name == "se100" else r2, K=K, trimmed="-T" in args)
    assert got == want
    # the same with two chunks in flight (chunk jobs) and the chunk's text as one buffer
    assert a.align_pipelined(r1, None if name == "se100" else r2, K=K, trimmed="-T" in args) == want
    # chaining stage: B200_CHAIN=check runs the device chaining AND the host chaining and aborts on any difference in the
    # chain/seed tables handed to chain2aln; B200_CHAIN=host is the host path alone (the one long reads take)
    for mode in ("check", "host"):
        os.environ["B200_CHAIN"] = mode
        try:
            got_m = a.align(r1, None if name == "se100" else r2, K=K, trimmed="-T" in args)
        finally:
            os.environ.pop("B200_CHAIN", None)
        assert got_m == want, mode
    # CIGAR stage: row windows sized too small on purpose, so that every region comes back flagged and takes the rerun pass
    os.environ["B200_GLOBAL_SQUEEZE"] = "1"
    try:
        got_sq = a.align(r1, None if name == "se100" else r2, K=K, trimmed="-T" in args)
    finally:
        os.environ.pop("B200_GLOBAL_SQUEEZE", None)
    assert got_sq == want
    # chunk jobs (b200_align_chunk_begin / _end): several chunks in flight, SAM unchanged and in input order
    assert a.align_pipelined(r1, None if name == "se100" else r2, K=K, trimmed="-T" in args) == want
    # from the raw fastq bytes, parsed by the job thread (one chunk: compare with a -K that does not split the input)
    if "-T" not in args:
        big = 1 << 40
        want1 = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "-t", "16", "-K", str(big), prefix] + fq, capture_output=True, check=True).stdout
        assert a.align_fastq(r1, None if name == "se100" else r2) == want1
    # through the stand-alone driver binary as well (the C host path), synchronous and with chunk jobs (-P)
    got2 = subprocess.run([os.path.join(ROOT, "tools", "b200_driver"), "-t", "16"] + args + [prefix] + fq, capture_output=True, check=True).stdout
    assert got2 == want
    got3 = subprocess.run([os.path.join(ROOT, "tools", "b200_driver"), "-P", "-t", "16"] + args + [prefix] + fq, capture_output=True, check=True).stdout
    assert got3 == want


def test_per_chromosome_routing(tmp_path):
    """row f3: the line table and the per-destination text of b200_set_routing / b200_align_chunk_end_routed (the routing kernels of
    finish_stage.h) against the ByChr host's rule applied to the plain SAM; chimeric pairs over five contigs fill 'discordant'"""
    import routing_check
    from mpibwa_b200 import simulate, index_build
    names, lengths, codes = simulate.make_reference(2_000_000, 5, seed=23)
    prefix = str(tmp_path / "ref.fa")
    index_build.write_fasta(prefix, simulate.codes_to_fasta_contigs(names, lengths, codes, n_runs=2))
    index_build.build_index(prefix)
    r1, r2 = simulate.simulate_pairs(codes, lengths, n_pairs=12000, unmappable_frac=0.05, seed=24)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    open(f1, "wb").write(r1); open(f2, "wb").write(r2)
    routing_check.make_chimeric(f2)
    drv = os.path.join(ROOT, "tools", "b200_driver")
    assert routing_check.check_driver(drv, ["-t", "8", "-K", "1200000", prefix, f1, f2], prefix) > 1000
    assert routing_check.check_driver(drv, ["-t", "8", "-K", "700000", prefix, f1], prefix) == 0      # single-end: no discordant


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref did not travel")
def test_odd_reads_vs_compiled_reference(tmp_path):
    """reads the simulator never makes - shorter than a seed, all N, homopolymers, exact copies of the reference, mates
    of very different lengths, reads long enough for mem_flt_chained_seeds (which moves the batch to the host chaining) -
    mixed into ordinary pairs: SAM == the compiled reference's, with blocking calls and with chunk jobs"""
    from mpibwa_b200 import simulate, index_build
    names, lengths, codes = simulate.make_reference(1_500_000, 3, seed=71)
    prefix = str(tmp_path / "ref.fa")
    index_build.build_index_from_codes(prefix, names, lengths, codes)
    rng = np.random.default_rng(5)
    B = np.frombuffer(b"ACGTN", np.uint8)

    def ref_piece(L, rc=False):
        s0 = int(rng.integers(1000, int(lengths[0]) - L - 1000))
        c = codes[s0:s0 + L]
        return B[(3 - c[::-1]) if rc else c].tobytes()

    odd = [(b"A", b"C"), (b"ACGTA", ref_piece(150)), (ref_piece(18), ref_piece(19)), (ref_piece(20, True), ref_piece(21)),
           (b"N" * 150, ref_piece(150)), (b"A" * 150, b"T" * 150), (ref_piece(150), ref_piece(150, True)),
           (ref_piece(75) + b"NNNNN" + ref_piece(70), ref_piece(150)), (ref_piece(150), ref_piece(35)),
           (ref_piece(260), ref_piece(300, True)), (b"ACGT" * 40, b"AC" * 75), (ref_piece(100) + ref_piece(100, True), ref_piece(150))]

    def block(pairs, tag):
        a, b = [], []
        for k, (x, y) in enumerate(pairs):
            a.append(b"@%s%04d/1\n%s\n+\n%s\n" % (tag, k, x, b"F" * len(x)))
            b.append(b"@%s%04d/2\n%s\n+\n%s\n" % (tag, k, y, b"F" * len(y)))
        return b"".join(a), b"".join(b)

    n1, n2 = simulate.simulate_pairs(codes, lengths, 3000, seed=9)
    o1, o2 = block(odd, b"odd")
    l1, l2 = simulate.simulate_pairs(codes, lengths, 24, read_len=800, ins_mean=1500, ins_sd=100, seed=10, prefix="long")
    m1, m2 = simulate.simulate_pairs(codes, lengths, 1000, seed=11, prefix="tail")
    for name, (r1, r2) in {"short": (n1 + o1 + m1, n2 + o2 + m2), "long": (n1 + l1 + o1 + m1, n2 + l2 + o2 + m2)}.items():
        f1, f2 = str(tmp_path / (name + "_1.fq")), str(tmp_path / (name + "_2.fq"))
        open(f1, "wb").write(r1); open(f2, "wb").write(r2)
        want = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "-t", "16", "-T", "-K", "400000", prefix, f1, f2],
                              capture_output=True, check=True).stdout
        a = M.Aligner(prefix, device=0, n_threads=16, verbose=1)
        assert a.align(r1, r2, K=400000, trimmed=True) == want, name
        assert a.align_pipelined(r1, r2, K=400000, trimmed=True) == want, name
        assert want.count(b"\n") >= 2 * (4000 + len(odd))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("opts", range(6))
def test_non_default_options_vs_compiled_reference(tmp_path, opts):
    """the option sets of tests/test_host_pipeline.py::OPTION_SETS through the real kernels (stand-alone driver, blocking and
    chunk jobs), indel-heavy reads, device chaining cross-checked against the host chaining"""
    from test_host_pipeline import OPTION_SETS, synthetic_case
    prefix, f1, f2 = synthetic_case(tmp_path, 12000, seed=13)
    args = ["-K", "1500000", "-o", OPTION_SETS[opts], prefix, f1, f2]
    want = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "-t", "16"] + args, capture_output=True, check=True).stdout
    drv = os.path.join(ROOT, "tools", "b200_driver")
    got = subprocess.run([drv, "-t", "16"] + args, capture_output=True, check=True, env=dict(os.environ, B200_CHAIN="check")).stdout
    assert got == want and want.count(b"\n") >= 24000
    assert subprocess.run([drv, "-P", "-t", "16"] + args, capture_output=True, check=True).stdout == want


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref did not travel")
def test_alt_contigs_vs_compiled_reference(tmp_path):
    """ALT-aware mapping (reference with a contig listed in <prefix>.alt) through the real kernels, blocking and chunk jobs"""
    from test_host_pipeline import alt_case
    prefix, f1, f2 = alt_case(tmp_path, 8000)
    args = ["-K", "1500000", prefix, f1, f2]
    want = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "-t", "16"] + args, capture_output=True, check=True).stdout
    drv = os.path.join(ROOT, "tools", "b200_driver")
    got = subprocess.run([drv, "-t", "16"] + args, capture_output=True, check=True, env=dict(os.environ, B200_CHAIN="check")).stdout
    assert got == want and want.count(b"pa:f:") > 2000
    assert subprocess.run([drv, "-P", "-t", "16"] + args, capture_output=True, check=True).stdout == want
    a = M.Aligner(prefix, device=0, n_threads=16, verbose=1)
    assert a.align(open(f1, "rb").read(), open(f2, "rb").read(), K=1500000) == want


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref did not travel")
def test_split_alignments_and_flags_vs_compiled_reference(tmp_path):
    """reads glued from two loci under every output flag (supplementary records, SA tags, NO_MULTI, PRIMARY5, KEEP_SUPP_MAPQ, ALL,
    SOFTCLIP, NOPAIRING) through the real kernels"""
    from test_host_pipeline import chimeric_case, FLAG_SETS
    prefix, f1, f2 = chimeric_case(tmp_path, 4000)
    drv = os.path.join(ROOT, "tools", "b200_driver")
    for flag in [0] + FLAG_SETS:
        args = ["-K", "900000", "-o", "flag=%d" % flag, prefix, f1, f2]
        want = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "-t", "16"] + args, capture_output=True, check=True).stdout
        got = subprocess.run([drv, "-P", "-t", "16"] + args, capture_output=True, check=True).stdout
        assert got == want, flag
        assert want.count(b"SA:Z:") > 5000 or flag == 0x4
    for pes in ("300,10,330,270", "450,30,520,380"):          # caller-given insert-size distribution (pes0 != NULL)
        args = ["-K", "900000", "-I", pes, prefix, f1, f2]
        want = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "-t", "16"] + args, capture_output=True, check=True).stdout
        assert subprocess.run([drv, "-t", "16"] + args, capture_output=True, check=True).stdout == want, pes


def test_properties_at_scale(tmp_path):
    """size-independent properties on a larger run: thread-count invariance, idempotence, record accounting"""
    from mpibwa_b200 import simulate, index_build
    names, lengths, codes = simulate.make_reference(20_000_000, 4, seed=23)
    prefix = str(tmp_path / "big.fa")
    index_build.build_index_from_codes(prefix, names, lengths, codes)
    r1, r2 = simulate.simulate_pairs(codes, lengths, 200000, seed=29)
    a = M.Aligner(prefix, device=0, n_threads=16, verbose=1)
    s16 = a.align(r1, r2, K=30_000_000)
    again = a.align(r1, r2, K=30_000_000)
    a.opt.contents.n_threads = 3
    s3 = a.align(r1, r2, K=30_000_000)
    os.environ["B200_CHAIN"] = "check"         # device chaining cross-checked against the host chaining on every read
    try:
        s_chk = a.align(r1, r2, K=30_000_000)
    finally:
        os.environ.pop("B200_CHAIN", None)
    assert s16 == again == s3 == s_chk
    # kernel-isolated replay of the chunk's ksw_extend2 job list (bench.py's one-batch figure): same cell count per job as the rounds
    os.environ["B200_EXT_RECORD"] = "1"
    try:
        assert a.align(r1, r2, K=1 << 40) == a.align_fastq(r1, r2)
        st = a.stats()
    finally:
        os.environ.pop("B200_EXT_RECORD", None)
    cells, jobs = C.c_int64(), C.c_int64()
    ms = a.lib.b200_ext_replay(a.opt, C.byref(cells), C.byref(jobs))
    assert ms > 0 and 0.9 * st["n_extend_jobs"] <= jobs.value <= st["n_extend_jobs"]
    assert 0.9 * st["extend_cells"] <= cells.value <= st["extend_cells"]
    flags = np.array([int(l.split(b"\t", 2)[1]) for l in s16.split(b"\n") if l])
    assert int(((flags & 0x900) == 0).sum()) == 400000            # one primary record per read
    assert ((flags & 4) == 0).mean() > 0.98                        # simulated reads map
