"""mpibwa_b200 - Python host-side mirror of the C ABI in include/mpibwa_b200.h (ctypes, no torch types).

The product is mpibwa_b200/libmpibwa_b200.so (C-ABI + host orchestration + sm_100a kernels).  This module only
declares the structs and prototypes of include/mpibwa_b200.h so that tests and bench.py can call the library the
way the mpiBWA hosts do.  There is no Python or CPU implementation of any compute entry point here: if the
library is missing, `load()` raises, and the library itself aborts when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libmpibwa_b200.so")     # (B200_LIB: A/B runs of two builds in one process environment)


class bwt_t(C.Structure):                     # reference src/bwt.h:46-58
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
                ("bwt", C.POINTER(C.c_uint32)), ("cnt_table", C.c_uint32 * 256), ("sa_intv", C.c_int),
                ("n_sa", C.c_uint64), ("sa", C.POINTER(C.c_uint64))]


class bwtintv_t(C.Structure):                 # reference src/bwt.h:60-62
    _fields_ = [("x", C.c_uint64 * 3), ("info", C.c_uint64)]


class bntann1_t(C.Structure):                 # reference src/bntseq.h:44-51
    _fields_ = [("offset", C.c_int64), ("len", C.c_int32), ("n_ambs", C.c_int32), ("gi", C.c_uint32),
                ("is_alt", C.c_int32), ("name", C.c_char_p), ("anno", C.c_char_p)]


class bntseq_t(C.Structure):                  # reference src/bntseq.h:59-67
    _fields_ = [("l_pac", C.c_int64), ("n_seqs", C.c_int32), ("seed", C.c_uint32), ("anns", C.POINTER(bntann1_t)),
                ("n_holes", C.c_int32), ("ambs", C.c_void_p), ("fp_pac", C.c_void_p)]


class bwaidx_t(C.Structure):                  # reference src/bwa.h:20-28
    _fields_ = [("bwt", C.POINTER(bwt_t)), ("bns", C.POINTER(bntseq_t)), ("pac", C.POINTER(C.c_uint8)),
                ("is_shm", C.c_int), ("l_mem", C.c_int64), ("mem", C.POINTER(C.c_uint8))]


class bseq1_t(C.Structure):                   # reference src/bwa.h:30-33
    _fields_ = [("l_seq", C.c_int), ("id", C.c_int), ("name", C.c_char_p), ("comment", C.c_char_p),
                ("seq", C.c_void_p), ("qual", C.c_char_p), ("sam", C.c_void_p)]


class mem_opt_t(C.Structure):                 # reference src/bwamem.h:25-57
    _fields_ = [("a", C.c_int), ("b", C.c_int), ("o_del", C.c_int), ("e_del", C.c_int), ("o_ins", C.c_int),
                ("e_ins", C.c_int), ("pen_unpaired", C.c_int), ("pen_clip5", C.c_int), ("pen_clip3", C.c_int),
                ("w", C.c_int), ("zdrop", C.c_int), ("max_mem_intv", C.c_uint64), ("T", C.c_int), ("flag", C.c_int),
                ("min_seed_len", C.c_int), ("min_chain_weight", C.c_int), ("max_chain_extend", C.c_int),
                ("split_factor", C.c_float), ("split_width", C.c_int), ("max_occ", C.c_int), ("max_chain_gap", C.c_int),
                ("n_threads", C.c_int), ("chunk_size", C.c_int), ("mask_level", C.c_float), ("drop_ratio", C.c_float),
                ("XA_drop_ratio", C.c_float), ("mask_level_redun", C.c_float), ("mapQ_coef_len", C.c_float),
                ("mapQ_coef_fac", C.c_int), ("max_ins", C.c_int), ("max_matesw", C.c_int), ("max_XA_hits", C.c_int),
                ("max_XA_hits_alt", C.c_int), ("mat", C.c_int8 * 25)]


class kswr_t(C.Structure):                    # reference src/ksw.h:14-19
    _fields_ = [("score", C.c_int), ("te", C.c_int), ("qe", C.c_int), ("score2", C.c_int), ("te2", C.c_int),
                ("tb", C.c_int), ("qb", C.c_int)]


class b200_extend_job_t(C.Structure):
    _fields_ = [("qlen", C.c_int32), ("tlen", C.c_int32), ("q_off", C.c_int64), ("t_off", C.c_int64),
                ("h0", C.c_int32), ("w", C.c_int32), ("end_bonus", C.c_int32),
                ("score", C.c_int32), ("qle", C.c_int32), ("tle", C.c_int32), ("gtle", C.c_int32),
                ("gscore", C.c_int32), ("max_off", C.c_int32)]


class b200_align_job_t(C.Structure):
    _fields_ = [("qlen", C.c_int32), ("tlen", C.c_int32), ("q_off", C.c_int64), ("t_off", C.c_int64),
                ("xtra", C.c_int32), ("r", kswr_t)]


class b200_global_job_t(C.Structure):
    _fields_ = [("qlen", C.c_int32), ("tlen", C.c_int32), ("q_off", C.c_int64), ("t_off", C.c_int64), ("w", C.c_int32),
                ("score", C.c_int32), ("n_cigar", C.c_int32), ("cigar_off", C.c_int64)]


class b200_stats_t(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ms_total", "ms_seed", "ms_sa", "ms_chain_host", "ms_extend", "ms_regs_host",
                                          "ms_rescue", "ms_sam_host", "ms_k_smem", "ms_k_sa", "ms_k_extend", "ms_k_sw",
                                          "ms_k_global")] + \
               [(n, C.c_int64) for n in ("n_reads", "n_bases", "n_intv", "n_seeds", "n_chains", "n_extend_jobs",
                                         "extend_cells", "n_sw_jobs", "sw_cells", "n_global_jobs", "global_cells",
                                         "fm_occ_blocks", "fm_sa_steps", "fm_sa_lookups", "n_launches", "h2d_bytes",
                                         "d2h_bytes")] + [("ms_k_extend_dp", C.c_double), ("n_extend_rounds", C.c_int64),
                                                          ("ms_sam_plan", C.c_double), ("ms_global", C.c_double),
                                                          ("n_global_host", C.c_int64), ("ms_k_chain", C.c_double),
                                                          ("ms_k_finish", C.c_double), ("ms_k_samtext", C.c_double),
                                                          ("ms_upload", C.c_double), ("ms_deliver", C.c_double)] + \
               [(n, C.c_int64) for n in ("n_patch_reads", "n_rescue_rounds", "n_global_rerun", "n_aln_slots", "sam_bytes")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


MEM_F_PE = 0x2
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000

# every symbol include/mpibwa_b200.h declares: (restype, argtypes)
_PROTOTYPES = {
    "mem_opt_init": (C.POINTER(mem_opt_t), []),
    "bwa_fill_scmat": (None, [C.c_int, C.c_int, C.POINTER(C.c_int8)]),
    "bwa_set_rg": (C.c_void_p, [C.c_char_p]),
    "bwa_insert_header": (C.c_void_p, [C.c_char_p, C.c_void_p]),
    "bwa_mem2idx": (C.c_int, [C.c_int64, C.POINTER(C.c_uint8), C.POINTER(bwaidx_t)]),
    "bwa_idx2mem": (C.c_int, [C.POINTER(bwaidx_t)]),
    "bwa_idx_load": (C.POINTER(bwaidx_t), [C.c_char_p, C.c_int]),
    "bwa_idx_destroy": (None, [C.POINTER(bwaidx_t)]),
    "mem_process_seqs": (None, [C.POINTER(mem_opt_t), C.POINTER(bwt_t), C.POINTER(bntseq_t), C.POINTER(C.c_uint8),
                                C.c_int64, C.c_int, C.POINTER(bseq1_t), C.c_void_p]),
    "ksw_extend2": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int8)] + [C.c_int] * 8 +
                    [C.POINTER(C.c_int)] * 5),
    "ksw_align2": (kswr_t, [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int8)] + [C.c_int] * 5 + [C.c_void_p]),
    "ksw_global2": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int8)] + [C.c_int] * 5 +
                    [C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32))]),
    "bwt_extend": (None, [C.POINTER(bwt_t), C.POINTER(bwtintv_t), C.POINTER(bwtintv_t), C.c_int]),
    "bwt_smem1": (C.c_int, [C.POINTER(bwt_t), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bwt_sa": (C.c_uint64, [C.POINTER(bwt_t), C.c_uint64]),
    "mem_chain2aln": (None, [C.POINTER(mem_opt_t), C.POINTER(bntseq_t), C.POINTER(C.c_uint8), C.c_int, C.c_char_p,
                             C.c_void_p, C.c_void_p]),
    "b200_gpu_init": (C.c_int, [C.POINTER(bwaidx_t), C.c_int]),
    "b200_gpu_release": (None, []),
    "b200_device_count": (C.c_int, []),
    "b200_ksw_extend2_batch": (C.c_int, [C.c_int64, C.POINTER(b200_extend_job_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.POINTER(C.c_int8)] + [C.c_int] * 5),
    "b200_ksw_align2_batch": (C.c_int, [C.c_int64, C.POINTER(b200_align_job_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                        C.POINTER(C.c_int8)] + [C.c_int] * 4),
    "b200_ksw_global2_batch": (C.c_int, [C.c_int64, C.POINTER(b200_global_job_t), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.POINTER(C.c_int8)] + [C.c_int] * 4 + [C.POINTER(C.POINTER(C.c_uint32))]),
    "b200_collect_intv_batch": (C.c_int, [C.POINTER(mem_opt_t), C.c_int, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.POINTER(bwtintv_t)), C.POINTER(C.POINTER(C.c_int64))]),
    "b200_bwt_sa_batch": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p]),
    "b200_fastq_parse": (C.c_int64, [C.c_void_p, C.c_int64, C.POINTER(C.POINTER(bseq1_t))]),
    "b200_set_host_threads": (None, [C.c_int]),
    "b200_plan_chunks": (C.c_int64, [C.c_int64, C.POINTER(bseq1_t), C.POINTER(bseq1_t), C.c_int64, C.c_int,
                                     C.POINTER(C.POINTER(C.c_int64))]),
    "b200_chunk_seqs": (C.POINTER(bseq1_t), [C.c_int64, C.POINTER(bseq1_t), C.POINTER(bseq1_t)]),
    "b200_collect_sam": (C.c_int64, [C.c_int64, C.POINTER(bseq1_t), C.POINTER(C.c_void_p)]),
    "b200_stage_reads": (None, [C.POINTER(mem_opt_t), C.POINTER(bwaidx_t), C.c_int, C.POINTER(bseq1_t)]),
    "b200_align_chunk": (C.c_int64, [C.POINTER(mem_opt_t), C.POINTER(bwaidx_t), C.c_int64, C.c_int64, C.POINTER(bseq1_t),
                                     C.POINTER(bseq1_t), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "b200_process_seqs_begin": (C.c_void_p, [C.POINTER(mem_opt_t), C.POINTER(bwt_t), C.POINTER(bntseq_t), C.POINTER(C.c_uint8),
                                             C.c_int64, C.c_int, C.POINTER(bseq1_t), C.c_void_p]),
    "b200_process_seqs_end": (None, [C.c_void_p, C.POINTER(b200_stats_t)]),
    "b200_align_chunk_begin": (C.c_void_p, [C.POINTER(mem_opt_t), C.POINTER(bwaidx_t), C.c_int64, C.c_int64, C.POINTER(bseq1_t),
                                            C.POINTER(bseq1_t)]),
    "b200_align_seqs_begin": (C.c_void_p, [C.POINTER(mem_opt_t), C.POINTER(bwaidx_t), C.c_int64, C.c_int, C.POINTER(bseq1_t), C.c_void_p]),
    "b200_align_fastq_begin": (C.c_void_p, [C.POINTER(mem_opt_t), C.POINTER(bwaidx_t), C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "b200_align_chunk_end": (C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(b200_stats_t)]),
    "b200_set_routing": (None, [C.c_int]),
    "b200_align_chunk_end_routed": (C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(b200_stats_t)]),
    "b200_free": (None, [C.c_void_p]),
    "b200_big_alloc": (C.c_void_p, [C.c_size_t]),
    "b200_get_stats": (None, [C.POINTER(b200_stats_t)]),
    "b200_get_aux_stats": (None, [C.POINTER(b200_stats_t)]),
    "b200_ext_replay": (C.c_double, [C.POINTER(mem_opt_t), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b200_hbm_random_sector_peak": (C.c_double, [C.c_int, C.c_size_t]),
    "b200_int32_peak": (C.c_double, [C.c_int]),
    "b200_int32_peak_dual_pipe": (C.c_double, [C.c_int]),
    "b200_version": (C.c_char_p, []),
}
_DATA_SYMBOLS = ("bwa_verbose", "bwa_rg_id", "bwa_pg")
EXPORTED_SYMBOLS = tuple(_PROTOTYPES) + _DATA_SYMBOLS

_lib = None


def build(verbose=False):
    """Compile the library in-tree (nvcc, sm_100a).  Returns the path of the .so."""
    r = subprocess.run(["make", "-C", ROOT, "lib", "driver"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libmpibwa_b200.so failed:\n" + (r.stdout or "")[-4000:] + (r.stderr or "")[-4000:])
    return LIB_PATH


def load(path=None):
    """dlopen the library and attach the prototypes.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("%s not found: run `make lib` (or __graft_entry__.build()); there is no Python/CPU fallback" % p)
    lib = C.CDLL(p)
    for name, (res, args) in _PROTOTYPES.items():
        if path is not None and not hasattr(lib, name):
            continue                      # test scaffolds (tests/hostemu) do not carry the device-only symbols
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class Aligner:
    """Stand-in for the mpiBWA chunk loop (reference src/mainParallel.c:1146-1493): load an index, parse fastq
    buffers in place, plan chunks with the reference's rule, call mem_process_seqs per chunk."""

    def __init__(self, index_prefix, device=0, n_threads=None, lib=None, paired=True, verbose=1):
        self.lib = lib or load()
        C.c_int.in_dll(self.lib, "bwa_verbose").value = verbose
        self.idx = self.lib.bwa_idx_load(index_prefix.encode(), 7)
        if not self.idx:
            raise RuntimeError("cannot load index " + index_prefix)
        self.opt = self.lib.mem_opt_init()
        self.opt.contents.n_threads = n_threads or os.cpu_count() or 1
        if paired:
            self.opt.contents.flag |= MEM_F_PE
        self.paired = paired
        self.lib.b200_gpu_init(self.idx, device)

    def parse(self, fastq_bytes):
        """-> (keepalive buffer, bseq1_t*, n).  The buffer is modified in place like the hosts do."""
        buf = C.create_string_buffer(fastq_bytes, len(fastq_bytes) + 1)
        seqs = C.POINTER(bseq1_t)()
        n = self.lib.b200_fastq_parse(C.cast(buf, C.c_void_p), len(fastq_bytes), C.byref(seqs))
        return buf, seqs, n

    def parse_array(self, arr, n_bytes):
        """Same on a caller-owned writable uint8 numpy array (one byte of slack after n_bytes): no extra copy."""
        seqs = C.POINTER(bseq1_t)()
        n = self.lib.b200_fastq_parse(C.c_void_p(arr.ctypes.data), n_bytes, C.byref(seqs))
        return arr, seqs, n

    def plan(self, n, s1, s2, K, trimmed=False):
        ends = C.POINTER(C.c_int64)()
        k = self.lib.b200_plan_chunks(n, s1, s2, K, int(trimmed), C.byref(ends))
        out = [ends[i] for i in range(k)]
        self.lib.b200_free(ends)
        return out

    def align_chunk(self, s1, s2, beg, end, n_processed=0, want_sam=True):
        sam = C.c_void_p()
        sam_len = C.c_int64()
        p1 = C.cast(C.addressof(s1.contents) + beg * C.sizeof(bseq1_t), C.POINTER(bseq1_t))
        p2 = C.cast(C.addressof(s2.contents) + beg * C.sizeof(bseq1_t), C.POINTER(bseq1_t)) if s2 else None
        total = self.lib.b200_align_chunk(self.opt, self.idx, n_processed, end - beg, p1, p2,
                                          C.byref(sam) if want_sam else None, C.byref(sam_len))
        out = None
        if want_sam:
            out = C.string_at(sam, sam_len.value)
            self.lib.b200_free(sam)
        return total, out

    def align(self, fq1, fq2=None, K=None, trimmed=False):
        """fastq bytes -> SAM bytes (records only), chunked like the hosts."""
        b1, s1, n1 = self.parse(fq1)
        b2, s2, n2 = (self.parse(fq2) if fq2 is not None else (None, None, n1))
        assert n1 == n2
        K = K or self.opt.contents.chunk_size * self.opt.contents.n_threads
        out, beg, n_proc = [], 0, 0
        for end in self.plan(n1, s1, s2, K, trimmed):
            total, sam = self.align_chunk(s1, s2, beg, end, n_proc if trimmed else 0)
            n_proc += total
            out.append(sam)
            beg = end
        self.lib.b200_free(s1)
        if s2:
            self.lib.b200_free(s2)
        return b"".join(out)

    def align_pipelined(self, fq1, fq2=None, K=None, trimmed=False):
        """align() with two chunks in flight (b200_align_chunk_begin / _end): same SAM, chunk i+1's device stages run under
        chunk i's host stages"""
        b1, s1, n1 = self.parse(fq1)
        b2, s2, n2 = (self.parse(fq2) if fq2 is not None else (None, None, n1))
        assert n1 == n2
        K = K or self.opt.contents.chunk_size * self.opt.contents.n_threads
        out, beg, n_proc, prev = [], 0, 0, None

        def finish(job):
            sam, sam_len = C.c_void_p(), C.c_int64()
            self.lib.b200_align_chunk_end(job, C.byref(sam), C.byref(sam_len), None)
            out.append(C.string_at(sam, sam_len.value))
            self.lib.b200_free(sam)

        for end in self.plan(n1, s1, s2, K, trimmed):
            p1 = C.cast(C.addressof(s1.contents) + beg * C.sizeof(bseq1_t), C.POINTER(bseq1_t))
            p2 = C.cast(C.addressof(s2.contents) + beg * C.sizeof(bseq1_t), C.POINTER(bseq1_t)) if s2 else None
            job = self.lib.b200_align_chunk_begin(self.opt, self.idx, n_proc if trimmed else 0, end - beg, p1, p2)
            n_proc += (end - beg) * (2 if s2 else 1)
            if prev is not None:
                finish(prev)
            prev, beg = job, end
        if prev is not None:
            finish(prev)
        self.lib.b200_free(s1)
        if s2:
            self.lib.b200_free(s2)
        return b"".join(out)

    def align_fastq(self, fq1, fq2=None):
        """the whole input as ONE chunk from its raw fastq bytes (b200_align_fastq_begin: parsed by the job thread)"""
        b1 = C.create_string_buffer(fq1, len(fq1) + 1)
        b2 = C.create_string_buffer(fq2, len(fq2) + 1) if fq2 is not None else None
        job = self.lib.b200_align_fastq_begin(self.opt, self.idx, 0, C.cast(b1, C.c_void_p), len(fq1),
                                              C.cast(b2, C.c_void_p) if b2 is not None else None, len(fq2) if fq2 is not None else 0)
        sam, sam_len = C.c_void_p(), C.c_int64()
        self.lib.b200_align_chunk_end(job, C.byref(sam), C.byref(sam_len), None)
        out = C.string_at(sam, sam_len.value)
        self.lib.b200_free(sam)
        return out

    def stats(self):
        st = b200_stats_t()
        self.lib.b200_get_stats(C.byref(st))
        return st.as_dict()

    def close(self):
        self.lib.b200_gpu_release()
