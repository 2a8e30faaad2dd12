import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fasttanhsinhquadrature.jl_b200"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): C restatement of the reference's scalar kernels."""
    from fts_oracle import Oracle, build
    build()
    return Oracle("strict")


@pytest.fixture(scope="session")
def fts():
    """The product: Python host mirror over libfts_b200.so (CUDA). No CPU fallback exists."""
    import fts_b200
    return fts_b200
