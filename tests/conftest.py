import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libssref.so")
B200_PLAN_LIB = os.path.join(ROOT, "supersonic_b200", "lib", "libssb200_plan.so")
B200_LIB = os.path.join(ROOT, "supersonic_b200", "lib", "libssb200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200; run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Builds (or finds) the in-tree libraries once per session."""
    import __graft_entry__ as entry
    entry.build()
    return True


@pytest.fixture(scope="session")
def ref(built):
    """The oracle: the unmodified reference behind the plan driver (test infrastructure)."""
    from supersonic_b200 import ssplan
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libssref.so missing (needs /root/reference to build)")
    return ssplan.PlanLib(REF_LIB)


@pytest.fixture(scope="session")
def b200(built):
    """The product: supersonic.h mirror + CUDA library behind the same plan driver."""
    from supersonic_b200 import ssplan
    return ssplan.PlanLib(B200_PLAN_LIB)
