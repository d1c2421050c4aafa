import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the reference's own unittest files, vendored as fixtures: run only through
# tests/test_reference_unittests_gpu.py (with `aesmc` aliased to this package), never collected directly
collect_ignore_glob = ["golden/ref_tests/*"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """tests/golden/reference_{default,avx2}.npz recorded from the unmodified reference."""
    import numpy as np
    here = os.path.join(ROOT, "tests", "golden")
    return {v: np.load(os.path.join(here, "reference_%s.npz" % v)) for v in ("default", "avx2")}


@pytest.fixture(scope="session")
def built():
    """Make sure libaesmc_b200.so and the oracle exist (builds them if a compiler is available)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def cuda(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
