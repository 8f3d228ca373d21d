import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `-m gpu`)")


@pytest.fixture(scope="session")
def bof():
    """The in-tree package (dlopens libbof_b200.so; raises if it was not built)."""
    import __graft_entry__ as g

    if not g.LIB.exists():
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def ctx(bof):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = bof.Context(device=0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle
