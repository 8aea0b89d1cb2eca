import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaugefields.jl_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import gf_oracle

    gf_oracle.build()
    return gf_oracle


@pytest.fixture(scope="session")
def backend():
    """One libgfb200 context per test session.  Fails loudly (no fallback) when no GPU is usable."""
    import gfb200

    return gfb200.B200Backend(ngpu=1)
