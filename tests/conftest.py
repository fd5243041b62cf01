import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference on seeded inputs (tests/golden/make_golden.py)."""
    with np.load(os.path.join(ROOT, "tests", "golden", "golden.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def vm():
    """The product package (loaded under the importable name ``video_matting_b200``)."""
    import __graft_entry__ as ge
    return ge.load_package()


def count_mismatch(a, b):
    return int(np.count_nonzero(np.asarray(a) != np.asarray(b)))
