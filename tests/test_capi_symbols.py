"""CPU test: the C-ABI library loads and exports every symbol that include/mpibwa_b200.h declares (no compute)."""
import ctypes as C
import os
import re
import pytest
from conftest import ROOT
import mpibwa_b200 as M


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mpibwa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", src))
    names |= set(re.findall(r"extern\s+[\w\s\*]+?\b(\w+)(?:\[\d+\])?\s*;", src))
    return names - {"defined", "sizeof"}


def test_header_and_binding_agree():
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    assert declared == set(M.EXPORTED_SYMBOLS), declared ^ set(M.EXPORTED_SYMBOLS)


def test_library_exports_every_symbol():
    if not os.path.exists(M.LIB_PATH):
        M.build()
    lib = C.CDLL(M.LIB_PATH)
    for name in M.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name
    M.load()
    assert b"mpibwa_b200" in M.load().b200_version()


def test_struct_layouts_match_reference_abi():
    # sizes of the reference structs on x86-64 LP64 (reference src/bwamem.h, src/bwt.h, src/bwa.h, src/bntseq.h)
    assert C.sizeof(M.bwt_t) == 1120
    assert C.sizeof(M.bwtintv_t) == 32
    assert C.sizeof(M.bseq1_t) == 48
    assert C.sizeof(M.bntann1_t) == 40
    assert C.sizeof(M.bntseq_t) == 48
    assert C.sizeof(M.bwaidx_t) == 48
    assert C.sizeof(M.kswr_t) == 28
    assert C.sizeof(M.mem_opt_t) == 168


def test_no_cpu_fallback_without_device():
    """On a machine without CUDA the product aborts loudly instead of computing on the CPU."""
    import subprocess
    import sys
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("CUDA device present")
    except ImportError:
        pass
    if not os.path.exists(M.LIB_PATH):
        M.build()
    code = ("import mpibwa_b200 as M, ctypes as C; L=M.load(); j=M.b200_extend_job_t(); "
            "L.b200_ksw_extend2_batch(1, C.byref(j), None, 0, None, 0, None, 6,1,6,1,100)")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0
    assert "no index on the device" in r.stderr or "no usable CUDA device" in r.stderr
