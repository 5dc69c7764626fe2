import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def stable_seed(*key) -> int:
    """Seed from a test's parameters that is the same in every process (hash() of a str is salted per process)."""
    import zlib
    return zlib.crc32(repr(key).encode())


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ssi():
    import subspaceinference_jl_b200 as mod
    return mod


@pytest.fixture()
def engine(ssi):
    eng = ssi.Engine(0)
    yield eng
    eng.close()
