import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts: compile the product library (nvcc, no GPU needed)
    and the oracle (gcc) once so that neither the ABI tests nor the GPU tests depend on a prior build()."""
    import importlib
    g = importlib.import_module("g2048_b200")
    if not os.path.exists(g.LIB_PATH):
        g.build_library()
    from oracle import pyoracle
    pyoracle.build()


@pytest.fixture(scope="session")
def golden():
    """Vectors generated from the live reference by oracle/make_golden.py."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_games():
    """Whole reference games at the BASELINE widths, move by move (oracle/make_golden_games.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_games.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle
