import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; built in-tree by `python -m uzu_b200.build` / __graft_entry__.build()."""
    from uzu_b200 import binding, build
    build.build()
    return binding.load()


@pytest.fixture(scope="session")
def ctx(lib):
    from uzu_b200 import binding
    c = binding.Context(0)      # raises without a GPU: there is no CPU fallback to fall back to
    yield c
    c.close()
