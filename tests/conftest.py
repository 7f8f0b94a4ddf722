import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_BUILD_ERROR = []


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts: compile the product library (nvcc, no GPU needed) and the oracle
    (gcc) once so that neither the ABI tests nor the GPU tests depend on a prior build().  A machine without nvcc
    still runs the tests that do not need the library (oracle, Philox, goldens): those that do are skipped."""
    import importlib
    g = importlib.import_module("g2048_b200")
    if not os.path.exists(g.LIB_PATH):
        try:
            g.build_library()
        except Exception as e:                  # noqa: BLE001 -- no nvcc here: remember why, skip what needs the library
            _BUILD_ERROR.append(str(e).splitlines()[0] if str(e) else repr(e))
    from oracle import pyoracle
    pyoracle.build()


def pytest_collection_modifyitems(config, items):
    if not _BUILD_ERROR:
        return
    skip = pytest.mark.skip(reason="libg2048.so could not be built here: " + _BUILD_ERROR[0])
    for item in items:
        if "gpu" in item.keywords or os.path.basename(str(item.fspath)) == "test_abi.py":
            item.add_marker(skip)


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
def golden_rollouts():
    """2,000-step random-policy rollouts of the live reference env, summarised per env (oracle/make_golden_rollouts.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "reference_rollouts.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle
