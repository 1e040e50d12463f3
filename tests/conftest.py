import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_artifacts_built():
    """The native artefacts are built in-tree by __graft_entry__.build(); build them if a fresh checkout has none
    (nvcc cross-compiles without a GPU).  This builds the product; it is not a fallback for it."""
    need = [os.path.join(ROOT, "kmers", "reps_b200", f) for f in ("libkmersreps_cuda.so", "libkrsynth.so", "host/kmers_reps", "host/selftest")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


def golden_bytes(name):
    p = os.path.join(GOLDEN, name)
    if os.path.exists(p + ".gz"):
        return gzip.decompress(open(p + ".gz", "rb").read())
    return open(p, "rb").read()


@pytest.fixture(scope="session")
def expected():
    return json.load(open(os.path.join(GOLDEN, "expected.json")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import kr_oracle
    return kr_oracle.get()


@pytest.fixture(scope="session")
def rep50():
    from oracle import kr_oracle as ko
    raw = golden_bytes("rep50.db")
    db = ko.parse_repdb(raw)
    db["raw"] = raw
    return db


@pytest.fixture(scope="session")
def vapb():
    from oracle import kr_oracle as ko
    return ko.read_fasta(golden_bytes("VapbProt.fa"))


@pytest.fixture(scope="session")
def gto_seeds():
    from oracle import kr_oracle as ko
    return ko.read_fasta(golden_bytes("gto_seeds.fa"))


@pytest.fixture(scope="session")
def rickettsia():
    rows = [ln.split("\t") for ln in golden_bytes("rickettsia_seed.tsv").decode().split("\n") if ln]
    return rows
