import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_arrays():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_arrays.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def synth():
    import avd_b200.synth as s
    return s


# tolerances (BASELINE.json north_star; SURVEY.md s8d "Parity gates")
FLOAT_RTOL = 1e-4
# absolute floor for Farneback magnitude statistics, in thumbnail pixels.  On static pairs the library's
# own optimized and non-optimized builds disagree by 2.9e-5 on mean(mag) (golden.json:
# cv2_self_disagreement_static) because the last-row/column branch of the update step is decided by the
# sign of a ~1e-6 flow.  Anything below this floor is not a property of the reference.
FLOW_ATOL = 5e-5


def close(a, b, rtol=FLOAT_RTOL, atol=1e-6):
    return abs(a - b) <= rtol * abs(b) + atol
