import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "easysimp.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def _has_gpu():
    try:
        import easysimp_b200 as es

        ctx = es.Context(0)
        ctx.close()
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    """Contexts for the -m gpu tests; there is no CPU fallback, so a missing device is an error."""
    import easysimp_b200 as es

    made = []

    def make():
        c = es.Context(0)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
