import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def decode_boundary(z):
    """Boundary spec stored by oracle/make_golden.py -> python object keyed by axis index."""
    vals = [str(v) for v in z["boundary_vals"]]

    def conv(s):
        if s in ("periodic", "reflect"):
            return s
        return float(s)

    if z["boundary_keys"].size == 0:
        return conv(vals[0])
    return {int(k): conv(v) for k, v in zip(z["boundary_keys"], vals)}


LAZY_CASES = ["x_periodic", "x_const0", "x_nan", "xy_periodic", "xy_mixed", "xy_one_axis", "xyz", "xy_land"]

# ESN kwargs every fixture was generated with (oracle/make_golden.py: ESN_KW)
ESN_KW = dict(
    input_kwargs=dict(factor=0.5, distribution="uniform", normalization="svd", random_seed=0),
    adjacency_kwargs=dict(factor=0.9, distribution="uniform", normalization="svd", is_sparse=True,
                          connectedness=5, format="csr", random_seed=1),
    bias_kwargs=dict(factor=0.5, distribution="uniform", random_seed=2),
)


@pytest.fixture(scope="session")
def esn_kw():
    return {k: dict(v) for k, v in ESN_KW.items()}
