"""CPU tests of the HOST orchestration (chaining, dedup/patch, insert-size statistics, rescue replay, pairing, mapQ,
CIGAR/MD, SAM text) against the reference's SAM.  The device stages are stood in for by tests/hostemu (the same
per-task bodies looped on the CPU) - a test scaffold only; the product library never contains it."""
import gzip
import hashlib
import json
import os
import subprocess
import pytest
from conftest import ROOT, have_ref

GOLD = os.path.join(ROOT, "tests", "golden")


def _head(path, n_reads, out):
    with open(path, "rb") as fi:
        lines = fi.read().split(b"\n")[:4 * n_reads]
    with open(out, "wb") as fo:
        fo.write(b"\n".join(lines) + b"\n")
    return out


def test_head1500_matches_golden_sam(hostemu_built, examples, tmp_path):
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    r1 = _head(examples["R1_10K"], 1500, str(tmp_path / "r1.fq"))
    r2 = _head(examples["R2_10K"], 1500, str(tmp_path / "r2.fq"))
    sam = subprocess.run([drv, "-t", "8", examples["idx"], r1, r2], capture_output=True, check=True).stdout
    with gzip.open(os.path.join(GOLD, "pe_head1500.sam.gz"), "rb") as fh:
        want = fh.read()
    assert sam == want
    assert hashlib.md5(sam).hexdigest() == json.load(open(os.path.join(GOLD, "sam_md5.json")))["pe_head1500"]["md5"]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", ["se_chunks", "trim", "threads"])
def test_shapes_match_reference(hostemu_built, examples, tmp_path, shape):
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if shape == "se_chunks":
        args = ["-K", "60000", examples["idx"], _head(examples["R1_10K"], 2000, str(tmp_path / "a.fq"))]
    elif shape == "trim":
        args = ["-T", "-K", "100000", examples["idx"], _head(examples["R1_10K_TRIM"], 1200, str(tmp_path / "a.fq")),
                _head(examples["R2_10K_TRIM"], 1200, str(tmp_path / "b.fq"))]
    else:
        args = ["-K", "150000", examples["idx"], _head(examples["R1_10K"], 1200, str(tmp_path / "a.fq")),
                _head(examples["R2_10K"], 1200, str(tmp_path / "b.fq"))]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    got = subprocess.run([drv, "-t", "1" if shape == "threads" else "8"] + args, capture_output=True, check=True).stdout
    assert got == want and want.count(b"\n") > 1000


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", ["pe", "trim"])
def test_chunk_jobs_give_the_same_sam(hostemu_built, examples, tmp_path, shape):
    """-P: the host loop keeps chunks in flight through b200_process_seqs_begin / _end (several chunks, each in its own slot of
    device buffers); records come out in input order and byte-identical to the reference's"""
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if shape == "pe":
        args = ["-K", "90000", examples["idx"], _head(examples["R1_10K"], 1500, str(tmp_path / "a.fq")),
                _head(examples["R2_10K"], 1500, str(tmp_path / "b.fq"))]
    else:
        args = ["-T", "-K", "70000", examples["idx"], _head(examples["R1_10K_TRIM"], 1200, str(tmp_path / "a.fq")),
                _head(examples["R2_10K_TRIM"], 1200, str(tmp_path / "b.fq"))]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    r = subprocess.run([drv, "-P", "-t", "4"] + args, capture_output=True, check=True)
    assert r.stdout == want and want.count(b"\n") > 2000
    # -C: b200_align_chunk (the chunk's text comes back as one buffer instead of seqs[i].sam)
    assert subprocess.run([drv, "-C", "-t", "4"] + args, capture_output=True, check=True).stdout == want
    # -F: b200_align_fastq_begin (the chunk goes in as raw fastq bytes: parsed, interleaved and encoded by the fastq stage)
    assert subprocess.run([drv, "-F", "-t", "4"] + args, capture_output=True, check=True).stdout == want


def test_occ_sectors_beyond_32_bits(hostemu_built):
    """the device's FM-index layout (32-byte occ sectors with 40-bit counts) against a direct count, at BWT rows of a
    human-sized reference (> 2^32) as well as small ones"""
    import ctypes as C
    lib = C.CDLL(os.path.join(hostemu_built, "libmpibwa_b200_hostemu.so"))
    lib.b200_emu_occ_selftest.restype = C.c_int64
    lib.b200_emu_occ_selftest.argtypes = [C.c_uint64, C.c_uint32]
    for k0 in (0, 128, 1 << 31, (1 << 32) - 128, 1 << 32, 6_200_000_000, (1 << 33) - 1024):
        for seed in (1, 2, 3):
            assert lib.b200_emu_occ_selftest(k0, seed) == 0, (k0, seed)
    # the packed interval entries of the seeding kernels hold three 33-bit values each
    lib.b200_emu_pack_selftest.restype = C.c_int64
    lib.b200_emu_pack_selftest.argtypes = [C.c_uint32]
    assert lib.b200_emu_pack_selftest(7) == 0


OPTION_SETS = ["w=200,zdrop=200",
               "a=2,b=5,o_del=7,e_del=2,o_ins=8,e_ins=1,pen_clip5=3,pen_clip3=7,pen_unpaired=10,T=40",
               "min_seed_len=15,min_chain_weight=10,max_occ=50,max_chain_extend=2,split_factor=1.2,split_width=5",
               "mask_level=0.3,drop_ratio=0.7,XA_drop_ratio=0.6,max_XA_hits=2,max_matesw=3,w=30,zdrop=50",
               "flag=32", "flag=128,max_chain_gap=100"]


def synthetic_case(tmp_path, n_pairs, seed=3):
    """small synthetic reference + indel-heavy pairs (band, z-drop and rescue paths fire); returns (prefix, r1 path, r2 path)"""
    import sys
    sys.path.insert(0, ROOT)
    from mpibwa_b200 import simulate, index_build
    names, lengths, codes = simulate.make_reference(1_200_000, 3, seed=seed)
    prefix = str(tmp_path / "ref.fa")
    index_build.build_index_from_codes(prefix, names, lengths, codes)
    r1, r2 = simulate.simulate_pairs(codes, lengths, n_pairs, sub=0.02, indel=0.006, max_indel=12, unmappable_frac=0.1, seed=seed + 1)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    open(f1, "wb").write(r1); open(f2, "wb").write(r2)
    return prefix, f1, f2


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("opts", OPTION_SETS[:4])
def test_non_default_options_match_reference(hostemu_built, tmp_path, opts):
    """alignment options away from their defaults (wide band / z-drop as in BASELINE configs[3], another scoring scheme with
    unequal gap costs, seeding and chaining thresholds, filter ratios): SAM == the compiled reference's with the same options;
    B200_CHAIN=check also compares the device chaining tables with the host chaining on the way"""
    prefix, f1, f2 = synthetic_case(tmp_path, 1200)
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    args = ["-K", "200000", "-o", opts, prefix, f1, f2]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    got = subprocess.run([drv, "-t", "4"] + args, capture_output=True, check=True, env=dict(os.environ, B200_CHAIN="check")).stdout
    assert got == want and want.count(b"\n") >= 2400


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_forced_finish_paths_match_reference(hostemu_built, tmp_path):
    """the rarely taken paths of the finish stages, forced: mate-rescue replays that miss a result and ask for one more job per
    round (B200_RESCUE_HIDE hides the last round-0 results of every pair) and CIGAR jobs whose band-doubling retry outgrows the
    row window of their class (B200_GLOBAL_SQUEEZE shrinks the windows)"""
    prefix, f1, f2 = synthetic_case(tmp_path, 1500, seed=11)
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    args = ["-K", "300000", prefix, f1, f2]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    r = subprocess.run([drv, "-t", "4", "-v", "3"] + args, capture_output=True, check=True,
                       env=dict(os.environ, B200_RESCUE_HIDE="2", B200_GLOBAL_SQUEEZE="1"))
    assert r.stdout == want
    assert b" 0 extra rescue rounds" not in r.stderr and b" 0 of " not in r.stderr


def alt_case(tmp_path, n_pairs):
    """reference with an ALT contig (a 2 % diverged copy of 60 kb of chr1, listed in <prefix>.alt) and reads drawn from both:
    ALT-aware chain filtering, primary marking, pa:f / XA / supplementary (SA) records"""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    from mpibwa_b200 import simulate, index_build
    names, lengths, codes = simulate.make_reference(1_000_000, 3, seed=31)
    rng = np.random.default_rng(7)
    seg = codes[100000:160000].copy()
    mut = rng.random(len(seg)) < 0.02
    seg[mut] = (seg[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    codes2 = np.concatenate([codes, seg])
    lengths2 = np.concatenate([lengths, [len(seg)]])
    prefix = str(tmp_path / "ref.fa")
    index_build.build_index_from_codes(prefix, names + ["chr1_alt1"], lengths2, codes2)
    open(prefix + ".alt", "w").write("@HD\tVN:1.0\nchr1_alt1\t0\tchr1\t100001\t60\t60000M\t*\t0\t0\t*\t*\n")
    r1, r2 = simulate.simulate_pairs(codes2, lengths2, n_pairs, seed=33)
    a1, a2 = simulate.simulate_pairs(np.concatenate([codes[90000:170000], seg]), np.array([80000, 60000]), n_pairs, seed=35, prefix="alt")
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    open(f1, "wb").write(r1 + a1); open(f2, "wb").write(r2 + a2)
    return prefix, f1, f2


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_alt_contigs_match_reference(hostemu_built, tmp_path):
    prefix, f1, f2 = alt_case(tmp_path, 1000)
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    args = ["-K", "400000", prefix, f1, f2]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    got = subprocess.run([drv, "-t", "4"] + args, capture_output=True, check=True, env=dict(os.environ, B200_CHAIN="check")).stdout
    assert got == want and want.count(b"pa:f:") > 300 and want.count(b"SA:Z:") > 300


def chimeric_case(tmp_path, n):
    """ordinary pairs + reads glued from two loci (split alignments: supplementary records, SA tags, the flags that govern them)"""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    from mpibwa_b200 import simulate, index_build
    names, lengths, codes = simulate.make_reference(800_000, 3, seed=41)
    prefix = str(tmp_path / "ref.fa")
    index_build.build_index_from_codes(prefix, names, lengths, codes)
    rng = np.random.default_rng(3)
    B = np.frombuffer(b"ACGTN", np.uint8)

    def piece(L, rc=False):
        s0 = int(rng.integers(1000, len(codes) - L - 1000))
        c = codes[s0:s0 + L]
        return B[(3 - c[::-1]) if rc else c].tobytes()

    a, b = [], []
    for k in range(n):
        l1 = int(rng.integers(40, 110))
        x = piece(l1, rng.random() < 0.5) + piece(150 - l1, rng.random() < 0.5)
        y = piece(150, rng.random() < 0.5) if rng.random() < 0.5 else piece(70) + piece(80, True)
        a.append(b"@chim%05d/1\n%s\n+\n%s\n" % (k, x, b"F" * 150))
        b.append(b"@chim%05d/2\n%s\n+\n%s\n" % (k, y, b"F" * 150))
    r1, r2 = simulate.simulate_pairs(codes, lengths, n, seed=43)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    open(f1, "wb").write(r1 + b"".join(a)); open(f2, "wb").write(r2 + b"".join(b))
    return prefix, f1, f2


FLAG_SETS = [0x10, 0x800, 0x1000, 0x810, 0x8, 0xA00, 0x4]   # NO_MULTI, PRIMARY5, KEEP_SUPP_MAPQ, both, ALL, SOFTCLIP|PRIMARY5, NOPAIRING


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flag", FLAG_SETS[:4])
def test_split_alignments_and_flags_match_reference(hostemu_built, tmp_path, flag):
    prefix, f1, f2 = chimeric_case(tmp_path, 600)
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    args = ["-K", "300000", "-o", "flag=%d" % flag, prefix, f1, f2]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    got = subprocess.run([drv, "-t", "4"] + args, capture_output=True, check=True).stdout
    assert got == want and want.count(b"SA:Z:") > 1000


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("pes", ["300,10,330,270", "450,30,520,380"])
def test_given_insert_size_distribution(hostemu_built, tmp_path, pes):
    """mem_process_seqs with pes0 != NULL (bwa mem -I): the chunk-global statistics are taken from the caller, rescue windows
    and pairing follow them"""
    prefix, f1, f2 = chimeric_case(tmp_path, 500)
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    args = ["-K", "300000", "-I", pes, prefix, f1, f2]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    assert subprocess.run([drv, "-t", "4"] + args, capture_output=True, check=True).stdout == want
    assert subprocess.run([drv, "-P", "-t", "4"] + args, capture_output=True, check=True).stdout == want



@pytest.mark.parametrize("shape", ["pe", "se", "chimeric"])
def test_per_chromosome_routing(hostemu_built, examples, tmp_path, shape):
    """b200_set_routing + b200_align_chunk_end_routed: the line table and the text grouped by destination against the reference
    host's own routing rule (tests/routing_check.py) applied to the plain SAM of the same chunks; 'chimeric': three contigs and
    every fifth pair with its second mate taken from another place, so that the discordant group fills"""
    import routing_check
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    if shape == "chimeric":
        prefix, f1, f2 = synthetic_case(tmp_path, 1000, seed=21)
        routing_check.make_chimeric(f2)
        args = ["-K", "80000", prefix, f1, f2]
    else:
        prefix = examples["idx"]
        args = ["-K", "90000", prefix, _head(examples["R1_10K"], 1200, str(tmp_path / "a.fq"))]
        if shape == "pe":
            args.append(_head(examples["R2_10K"], 1200, str(tmp_path / "b.fq")))
    n_disc = routing_check.check_driver(drv, ["-t", "4"] + args, prefix)
    assert (n_disc > 50) == (shape == "chimeric")


def test_derived_index_structures(hostemu_built, examples):
    """round 2's HBM-resident structures - k-mer interval tables, whole suffix array + inverse, Bloom filters over the text's 19-mers -
    built by the routines the upload kernels run (tests/hostemu) against the plain FM-index routines: table entries == iterated
    bwt_extend (forwards and backwards, absent patterns included), suffix array == the walk to a sampled row, filters without false
    negatives and with a false-positive rate of a few percent"""
    import ctypes as C
    import sys
    sys.path.insert(0, ROOT)
    import mpibwa_b200 as M
    lib = M.load(os.path.join(hostemu_built, "libmpibwa_b200_hostemu.so"))
    idx = lib.bwa_idx_load(examples["idx"].encode(), 7)
    assert idx
    fn = lib.b200_emu_index_tables_selftest
    fn.restype = C.c_int64
    fn.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    n, fp = C.c_int64(), C.c_double()
    assert fn(idx, 5, C.byref(n), C.byref(fp)) == 0
    assert n.value > 10000 and fp.value < 0.05


SEEDING_SETS = [("min_seed_len=10", {}), ("min_seed_len=25,split_factor=1.1", {}), ("min_seed_len=32,max_occ=20", {}),
                ("min_seed_len=19,split_width=2", {"B200_KMER_MAX": "12"}), ("min_seed_len=14", {"B200_KMER_MAX": "6", "B200_BLOOM": "0"}),
                ("min_seed_len=19", {"B200_SA_FULL": "1"}), ("min_seed_len=21", {"B200_KMER_MAX": "0", "B200_SA_FULL": "0"})]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", range(len(SEEDING_SETS)))
def test_seeding_paths_under_options(hostemu_built, tmp_path, case):
    """the table-driven seeding (k-mer tables, backward chains, Bloom filters, unique walks) under seed lengths below / at / above
    the tables' depth and the filters' window, with shallow or absent tables, on reads full of ambiguous bases, of 20 to 330 bases
    and with planted repeats: tests/hostemu cross-checks the sweeps against the plain restatement on EVERY read and aborts on a
    difference; the SAM must be the compiled reference's"""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    from mpibwa_b200 import simulate, index_build
    opts, env = SEEDING_SETS[case]
    names, lengths, codes = simulate.make_reference(400_000, 3, seed=40 + case)
    codes = codes.copy()
    codes[150_000:158_000] = codes[20_000:28_000]            # an exact 8 kb repeat: long shared SMEMs, intervals of size 2
    codes[300_000:301_500] = codes[21_000:22_500]            # ... and of size 3
    prefix = str(tmp_path / "ref.fa")
    index_build.build_index_from_codes(prefix, names, lengths, codes)
    r1, r2 = simulate.simulate_pairs(codes, lengths, 700, read_len=330, sub=0.02, indel=0.004, max_indel=6, n_rate=0.01,
                                     trim_to=(20, 330), unmappable_frac=0.05, seed=50 + case)
    f1, f2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    open(f1, "wb").write(r1); open(f2, "wb").write(r2)
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    args = ["-T", "-K", "150000", "-o", opts, prefix, f1, f2]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    r = subprocess.run([drv, "-t", "4"] + args, capture_output=True, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == want and want.count(b"\n") >= 1400


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("env", [{}, {"B200_KMER_MAX": "6"}, {"B200_BLOOM": "0", "B200_SA_FULL": "0"}])
def test_seeding_on_repeats(hostemu_built, tmp_path, env):
    """reads out of tandem repeats, a homopolymer run and sixty 1-3 % diverged copies of a 300-base unit: backward entries that
    survive together for many bases (the merge stand-in and the chain budget of the backward sweeps), huge intervals, long lists -
    cross-checked per read inside tests/hostemu, SAM == the compiled reference's"""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    from mpibwa_b200 import simulate, index_build
    rng = np.random.default_rng(5)
    names, lengths, codes = simulate.make_reference(300_000, 2, seed=91)
    codes = codes.copy()
    unit = codes[1000:1300].copy()
    for k in range(60):
        u = unit.copy()
        m = rng.random(300) < 0.02
        u[m] = rng.integers(0, 4, m.sum())
        codes[5000 + k * 2000:5300 + k * 2000] = u
    codes[150000:151000] = np.tile(np.array([0, 1], dtype=codes.dtype), 500)
    codes[152000:152600] = 0
    codes[153000:154200] = np.tile(codes[153000:153012], 100)
    prefix = str(tmp_path / "rep.fa")
    index_build.build_index_from_codes(prefix, names, lengths, codes)

    def fq(name, seq):
        return ("@%s\n%s\n+\n%s\n" % (name, "".join("ACGT"[c] for c in seq), "I" * len(seq))).encode()

    a = b = b""
    for i in range(400):
        z = [5000 + int(rng.integers(0, 60)) * 2000 + int(rng.integers(0, 150)), 150000 + int(rng.integers(0, 850)),
             152000 + int(rng.integers(0, 450)), 153000 + int(rng.integers(0, 1050))][i % 4]
        s = codes[z:z + 150].copy()
        m = rng.random(150) < 0.01
        s[m] = rng.integers(0, 4, m.sum())
        a += fq("rep%d" % i, s)
        b += fq("rep%d" % i, (3 - codes[z + 200:z + 350])[::-1])
    f1, f2 = str(tmp_path / "q1.fq"), str(tmp_path / "q2.fq")
    open(f1, "wb").write(a); open(f2, "wb").write(b)
    drv = os.path.join(hostemu_built, "b200_driver_hostemu")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    args = ["-K", "200000", prefix, f1, f2]
    want = subprocess.run([ref, "-t", "4"] + args, capture_output=True, check=True).stdout
    r = subprocess.run([drv, "-t", "4"] + args, capture_output=True, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == want and want.count(b"\n") >= 800
