import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pkg(sub: str = ""):
    """Import the package; its name starts with a digit, so `import` syntax cannot be used."""
    name = "3ddctvideoencoding_b200" + ("." + sub if sub else "")
    return importlib.import_module(name)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def synth():
    return pkg("synth")


@pytest.fixture(scope="session")
def kat():
    import json
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_golden.npz"))
