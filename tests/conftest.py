import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """A pk_ctx on cuda:0.  GPU tests fail (not skip) when the CUDA library cannot run: there is no fallback."""
    import piranha_b200

    c = piranha_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    from oracle import cpu_oracle

    cpu_oracle.build()
