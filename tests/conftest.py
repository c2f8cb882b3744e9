"""pytest configuration: the `gpu` marker, build-on-demand of the test-only artefacts, shared data fixtures."""
import os
import subprocess
import sys
import tarfile
import gzip
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _make(*targets):
    r = subprocess.run(["make", "-C", ROOT, *targets], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.fixture(scope="session")
def oracle_built():
    """liboracle.so (our C restatement) and, when the reference sources or a prebuilt copy exist, oracle/_ref."""
    _make("oracle")
    return os.path.join(ROOT, "oracle")


@pytest.fixture(scope="session")
def hostemu_built():
    _make("hostemu")
    return os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="session")
def examples(tmp_path_factory):
    """Unpacked copy of the reference's example data (tests/golden/examples): index prefix + fastq paths."""
    d = tmp_path_factory.mktemp("examples")
    src = os.path.join(ROOT, "tests", "golden", "examples")
    with tarfile.open(os.path.join(src, "hg19.small.tar.gz")) as tf:
        tf.extractall(d)
    fa = None
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(".fa"):
                fa = os.path.join(base, f)
    out = {"idx": fa}
    for key in ("R1_10K", "R2_10K", "R1_10K_TRIM", "R2_10K_TRIM"):
        p = os.path.join(d, key + ".fq")
        with gzip.open(os.path.join(src, "HCC1187C_%s.fastq.gz" % key), "rb") as fi, open(p, "wb") as fo:
            fo.write(fi.read())
        out[key] = p
    return out


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbwa_ref.so"))


def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
