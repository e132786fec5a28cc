import gzip
import pickle
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load(name):
    with gzip.open(GOLDEN / (name + ".gz"), "rb") as fh:
        return pickle.load(fh)


@pytest.fixture(scope="session")
def unit_vectors():
    return _load("unit_vectors.pkl")


@pytest.fixture(scope="session")
def small_cases():
    return _load("small_find_motif.pkl")


@pytest.fixture(scope="session")
def testfa():
    return _load("testfa.pkl")


@pytest.fixture(scope="session")
def testfa_stock_k():
    return _load("testfa_stock_k.pkl")


@pytest.fixture(scope="session")
def r2_vectors():
    return _load("r2_vectors.pkl")


@pytest.fixture(scope="session")
def motif_def_file():
    return str(ROOT / "kmap_b200" / "default_motif_def_table.csv")
