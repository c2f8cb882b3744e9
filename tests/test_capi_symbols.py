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


REF_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference sources not present (they do not travel to the GPU box)")
def test_c_header_layouts_equal_the_reference_headers(tmp_path):
    """include/mpibwa_b200.h against the reference's OWN headers: one C translation unit compiled against each prints sizeof and
    offsetof of every ABI type (and the flag constants); the outputs must be identical"""
    import subprocess
    src = os.path.join(ROOT, "tests", "abi", "abi_probe.c")
    cfg = tmp_path / "config.h"
    cfg.write_text('#define VERSION "1.5.5"\n')
    outs = []
    for name, flags in (("ref", ["-DUSE_REF", "-I" + REF_SRC, "-I" + str(tmp_path)]), ("b200", ["-I" + os.path.join(ROOT, "include")])):
        exe = str(tmp_path / ("probe_" + name))
        subprocess.run(["gcc", "-w", "-o", exe, src] + flags, check=True)
        outs.append(subprocess.run([exe], capture_output=True, check=True, text=True).stdout)
    assert outs[0] == outs[1] and outs[0].count("\n") > 100


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference sources not present (they do not travel to the GPU box)")
def test_reference_pidx_links_against_the_library(hostemu_built, examples, tmp_path):
    """the reference's own src/pidx.c (the mpiBWAIdx host: needs no MPI) compiled unmodified against the library's symbols writes
    the same .map image as the reference build of the same file (oracle/_ref/mpiBWAIdx)"""
    import shutil
    import subprocess
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "mpiBWAIdx")
    if not os.path.exists(ref_tool):
        pytest.skip("oracle/_ref/mpiBWAIdx not built")
    cfg = tmp_path / "config.h"
    cfg.write_text('#define VERSION "1.5.5"\n')
    exe = str(tmp_path / "mpiBWAIdx_b200")
    # (the CPU test scaffold carries the same capi.cpp as the product library; pidx.c touches only bwa_idx_load / bwa_idx2mem)
    subprocess.run(["gcc", "-O2", "-w", "-I" + REF_SRC, "-I" + str(tmp_path), os.path.join(REF_SRC, "pidx.c"), "-o", exe,
                    "-L" + hostemu_built, "-lmpibwa_b200_hostemu", "-Wl,-rpath," + hostemu_built], check=True)
    maps = []
    for tool, tag in ((exe, "b200"), (ref_tool, "ref")):
        d = tmp_path / tag
        d.mkdir()
        for ext in (".bwt", ".sa", ".pac", ".ann", ".amb"):
            shutil.copy(examples["idx"] + ext, str(d / ("x.fa" + ext)))
        subprocess.run([tool, str(d / "x.fa")], check=True, capture_output=True)
        maps.append(open(str(d / "x.fa.map"), "rb").read())
    # the image holds raw pointers (bwt->bwt, bwt->sa, anns[].name ...) that differ from run to run: compare with them masked
    import numpy as np
    a, b = (np.frombuffer(m, np.uint8).copy() for m in maps)
    assert len(a) == len(b) and len(a) > 1_000_000
    diff = np.flatnonzero(a != b)
    assert len(diff) < 64 * (2 + 8), len(diff)          # only the pointer fields of bwt_t, bntseq_t and the annotation records


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
