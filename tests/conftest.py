import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import inflexion_b200  # noqa: E402,F401  registers the package alias

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", params=["hostsim", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """'cuda' = the product library on the GPU (the parity tests proper, -m gpu);
    'hostsim' = the same kernel sources compiled for the CPU (logic check in the GPU-less container)."""
    if request.param == "cuda":
        from inflexion_b200.backend import Backend
        return Backend.cuda()
    from hostsim import hostsim_backend
    return hostsim_backend()


@pytest.fixture(scope="session")
def cuda_backend():
    from inflexion_b200.backend import Backend
    return Backend.cuda()


@pytest.hookimpl(trylast=True)
def pytest_collection_modifyitems(config, items):
    """Keep tests/test_zz_late_additions.py (cases not yet run on a GPU) at the very end, also after pytest has grouped
    the tests by the session-scoped ``backend`` parameter, so that ``-x`` cannot let them mask the rest of the suite."""
    late = [it for it in items if "test_zz_late_additions" in it.nodeid]
    if late:
        items[:] = [it for it in items if "test_zz_late_additions" not in it.nodeid] + late
