"""Shared fixtures.  `-m "not gpu"`: oracle vs golden vectors, host C surface, ABI exports, sharding logic
(gloo, world_size 2).  `-m gpu`: parity of the CUDA path against the oracle, always through the C ABI."""
import importlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def bs():
    import __graft_entry__ as ge
    ge.build(verbose=False)
    return importlib.import_module("binary-spgemm_b200")


@pytest.fixture(scope="session")
def oracle(bs):
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference; prebuilt files travel to the GPU box)")
    return Ref()


@pytest.fixture(scope="session")
def fixture_npz():
    return np.load(GOLDEN / "validity_fixture.npz")


@pytest.fixture(scope="session")
def kats():
    return json.loads((GOLDEN / "kats.json").read_text())


@pytest.fixture(scope="session")
def seeded_cases():
    return json.loads((GOLDEN / "seeded_cases.json").read_text())


@pytest.fixture(scope="session")
def gpu_ctx(bs):
    """Process-wide bspgemm context on GPU 0 (fails loudly when the library or the GPU is missing)."""
    bs.init(1)
    yield bs
    bs.finalize()
