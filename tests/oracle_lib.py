"""TEST INFRASTRUCTURE: ctypes access to oracle/liboracle.so (our C restatement) and to oracle/_ref/libbwa_ref.so
(the unmodified reference sources compiled by oracle/Makefile).  Only tests/, smoke() and bench.py's CPU baseline
import this module."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbwa_ref.so")


class orc_ext_t(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("score", "qle", "tle", "gtle", "gscore", "max_off")] + [("cells", C.c_int64)]


class orc_aln_t(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("score", "te", "qe", "score2", "te2", "tb", "qb")] + [("cells", C.c_int64)]


class orc_intv_t(C.Structure):
    _fields_ = [("x0", C.c_uint64), ("x1", C.c_uint64), ("x2", C.c_uint64), ("info", C.c_uint64)]


class orc_fm_t(C.Structure):
    _fields_ = [("bwt", C.c_void_p), ("sa", C.c_void_p), ("primary", C.c_uint64), ("L2", C.c_uint64 * 5),
                ("seq_len", C.c_uint64), ("sa_intv", C.c_int)]


def default_mat(a=1, b=4):
    m = []
    for i in range(4):
        m += [a if i == j else -b for j in range(4)] + [-1]
    m += [-1] * 5
    return (C.c_int8 * 25)(*m)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_ksw_extend2.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int8)] + [C.c_int] * 8 + [C.POINTER(orc_ext_t)]
        L.orc_ksw_align2.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int8)] + [C.c_int] * 5 + [C.POINTER(orc_aln_t)]
        L.orc_ksw_global2.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int8)] + [C.c_int] * 5 + \
            [C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_int64)]
        L.orc_ksw_global2.restype = C.c_int
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        L.orc_extend.argtypes = [C.POINTER(orc_fm_t), C.POINTER(orc_intv_t), C.POINTER(orc_intv_t), C.c_int]
        L.orc_smem1.argtypes = [C.POINTER(orc_fm_t), C.c_int, C.c_char_p, C.c_int, C.c_uint64, C.POINTER(orc_intv_t), C.POINTER(C.c_int)]
        L.orc_smem1.restype = C.c_int
        L.orc_seed_strategy1.argtypes = [C.POINTER(orc_fm_t), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(orc_intv_t)]
        L.orc_collect_intv.argtypes = [C.POINTER(orc_fm_t), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(orc_intv_t)]
        L.orc_sa.argtypes = [C.POINTER(orc_fm_t), C.c_uint64]
        L.orc_sa.restype = C.c_uint64
        L.orc_get_seq.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_get_seq.restype = C.c_int64

    def extend(self, q, t, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
        r = orc_ext_t()
        self.lib.orc_ksw_extend2(len(q), bytes(q), len(t), bytes(t), mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, C.byref(r))
        return (r.score, r.qle, r.tle, r.gtle, r.gscore, r.max_off), r.cells

    def align(self, q, t, mat, o_del, e_del, o_ins, e_ins, xtra):
        r = orc_aln_t()
        self.lib.orc_ksw_align2(len(q), bytes(q), len(t), bytes(t), mat, o_del, e_del, o_ins, e_ins, xtra, C.byref(r))
        return (r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb), r.cells

    def global2(self, q, t, mat, o_del, e_del, o_ins, e_ins, w):
        """-> (score, [cigar ops as len << 4 | op]), band cells"""
        n, cg, cells = C.c_int(), C.POINTER(C.c_uint32)(), C.c_int64()
        sc = self.lib.orc_ksw_global2(len(q), bytes(q), len(t), bytes(t), mat, o_del, e_del, o_ins, e_ins, w, C.byref(n), C.byref(cg), C.byref(cells))
        ops = [cg[i] for i in range(n.value)]
        self.libc.free(cg)
        return (sc, ops), cells.value

    def collect_intv(self, fm, seq, min_seed_len=19, split_factor=1.5, split_width=10, max_mem_intv=20):
        out = (orc_intv_t * (3 * len(seq) + 8))()
        n = self.lib.orc_collect_intv(C.byref(fm), min_seed_len, split_factor, split_width, max_mem_intv, len(seq), bytes(seq), out)
        return [(out[i].x0, out[i].x1, out[i].x2, out[i].info) for i in range(n)]


class IndexFiles:
    """Raw views of <prefix>.bwt/.sa/.pac for the oracle (formats: reference src/bwt.c:421-462, src/bntseq.c:224)."""

    def __init__(self, prefix):
        raw = np.fromfile(prefix + ".bwt", dtype=np.uint64, count=5)
        self.primary = int(raw[0])
        self.L2 = [0] + [int(v) for v in raw[1:5]]
        self.seq_len = self.L2[4]
        self.bwt = np.fromfile(prefix + ".bwt", dtype=np.uint32, offset=40)
        hdr = np.fromfile(prefix + ".sa", dtype=np.uint64, count=7)
        self.sa_intv = int(hdr[5])
        self.sa = np.concatenate([np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64), np.fromfile(prefix + ".sa", dtype=np.uint64, offset=56)])
        self.pac = np.fromfile(prefix + ".pac", dtype=np.uint8)
        self.l_pac = self.seq_len // 2
        self.fm = orc_fm_t(self.bwt.ctypes.data, self.sa.ctypes.data, self.primary, (C.c_uint64 * 5)(*self.L2), self.seq_len, self.sa_intv)


# ---- the compiled reference -------------------------------------------------------------------------------------

class kswr_t(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("score", "te", "qe", "score2", "te2", "tb", "qb")]


class ref_bwtintv_t(C.Structure):
    _fields_ = [("x", C.c_uint64 * 3), ("info", C.c_uint64)]


class ref_bwtintv_v(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(ref_bwtintv_t))]


class Reference:
    """The unmodified reference library.  Struct layouts come from mpibwa_b200 (they mirror the reference ABI)."""

    def __init__(self):
        import mpibwa_b200 as M
        self.M = M
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ksw_extend2.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int8)] + [C.c_int] * 8 + [C.POINTER(C.c_int)] * 5
        L.ksw_align2.restype = kswr_t
        L.ksw_align2.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int8)] + [C.c_int] * 5 + [C.c_void_p]
        L.ksw_global2.restype = C.c_int
        L.ksw_global2.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int8)] + [C.c_int] * 5 + \
            [C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32))]
        L.bwa_idx_load.restype = C.POINTER(M.bwaidx_t)
        L.bwa_idx_load.argtypes = [C.c_char_p, C.c_int]
        L.bwt_extend.argtypes = [C.POINTER(M.bwt_t), C.POINTER(ref_bwtintv_t), C.POINTER(ref_bwtintv_t), C.c_int]
        L.bwt_smem1.argtypes = [C.POINTER(M.bwt_t), C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(ref_bwtintv_v), C.c_void_p]
        L.bwt_seed_strategy1.argtypes = [C.POINTER(M.bwt_t), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(ref_bwtintv_t)]
        L.bwt_sa.restype = C.c_uint64
        L.bwt_sa.argtypes = [C.POINTER(M.bwt_t), C.c_uint64]
        L.bns_get_seq.restype = C.POINTER(C.c_uint8)
        L.bns_get_seq.argtypes = [C.c_int64, C.POINTER(C.c_uint8), C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    def extend(self, q, t, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
        o = [C.c_int() for _ in range(5)]
        sc = self.lib.ksw_extend2(len(q), bytes(q), len(t), bytes(t), 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                                  *[C.byref(x) for x in o])
        return (sc, o[0].value, o[1].value, o[2].value, o[3].value, o[4].value)

    def align(self, q, t, mat, o_del, e_del, o_ins, e_ins, xtra):
        qb, tb = C.create_string_buffer(bytes(q), len(q) + 16), C.create_string_buffer(bytes(t), len(t) + 16)
        r = self.lib.ksw_align2(len(q), qb, len(t), tb, 5, mat, o_del, e_del, o_ins, e_ins, xtra, None)
        return (r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb)

    def global2(self, q, t, mat, o_del, e_del, o_ins, e_ins, w):
        n, cg = C.c_int(), C.POINTER(C.c_uint32)()
        sc = self.lib.ksw_global2(len(q), bytes(q), len(t), bytes(t), 5, mat, o_del, e_del, o_ins, e_ins, w, C.byref(n), C.byref(cg))
        ops = [cg[i] for i in range(n.value)]
        self.libc.free(cg)
        return (sc, ops)

    def smem1(self, bwt, seq, x, min_intv):
        mem = ref_bwtintv_v()
        ret = self.lib.bwt_smem1(bwt, len(seq), bytes(seq), x, min_intv, C.byref(mem), None)
        out = [(mem.a[i].x[0], mem.a[i].x[1], mem.a[i].x[2], mem.a[i].info) for i in range(mem.n)]
        self.libc.free(mem.a)
        return ret, out
