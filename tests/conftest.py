import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name + ".npz"))


def fiber_kwargs(g):
    import numpy as np

    kw = {}
    for k, v in zip(g["kw_names"], g["kw_vals"]):
        kw[str(k)] = None if np.isnan(v) else float(v)
    return kw


@pytest.fixture(scope="session")
def have_reference():
    from oracle.ref_shim import reference_root

    return reference_root() is not None
