import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref():
    """The UNMODIFIED reference compiled into oracle/_ref (test infrastructure only)."""
    from tests.refshim import RefLib
    try:
        return RefLib()
    except OSError as e:  # pragma: no cover
        pytest.skip("oracle/_ref/libsecp256k1_ref.so not built: %s" % e)


@pytest.fixture(scope="session")
def engine():
    import secp256k1_zkp_b200 as b200
    eng = b200.Engine()
    yield eng
    eng.close()
