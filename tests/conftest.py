import lzma
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_BIN = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def build_index(fasta: str, prefix: str) -> None:
    """Index producer for tests: the unmodified reference indexer (oracle/_ref/bwt_index)."""
    exe = os.path.join(REF_BIN, "bwt_index")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bwt_index not built (run `make -C oracle ref` where /root/reference exists)")
    subprocess.run([exe, fasta, prefix], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("gsa"))


@pytest.fixture(scope="session")
def ecoli(workdir):
    """config C1: test/ecoli.fa vs test/ecoli.mut of the reference repo (committed as xz fixtures)."""
    d = os.path.join(workdir, "test")
    os.makedirs(d, exist_ok=True)
    for n in ("ecoli.fa", "ecoli.mut"):
        with lzma.open(os.path.join(GOLDEN, n + ".xz"), "rb") as f, open(os.path.join(d, n), "wb") as g:
            g.write(f.read())
    build_index(os.path.join(d, "ecoli.fa"), os.path.join(d, "ecoli"))
    from gsalign_b200 import bwaidx, synth
    return {"dir": d, "prefix": os.path.join(d, "ecoli"), "index": bwaidx.load(os.path.join(d, "ecoli")),
            "query": synth.read_fasta(os.path.join(d, "ecoli.mut"))[0][1].tobytes(), "query_path": os.path.join(d, "ecoli.mut")}


@pytest.fixture(scope="session")
def oracle():
    import orc
    return orc.Oracle()
