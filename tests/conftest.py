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
    """Inputs (fresh copies), keyword arguments and outputs of one golden case."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    T = int(z["T"])
    R = [z["R%d" % t].copy() for t in range(T)]
    L = [z["len%d" % t].copy() for t in range(T)]
    P = [z["prob%d" % t].copy() for t in range(T)]
    kwargs = {}
    for k in z.files:
        if k.startswith("arg_"):
            v = z[k]
            if v.ndim == 0:
                v = v.item()
                if isinstance(v, float) and np.isnan(v):
                    v = None
                elif k[4:] in ("M", "initial", "gap", "randomS"):
                    v = int(v)
            kwargs[k[4:]] = v
    return R, L, P, kwargs, z


@pytest.fixture
def golden():
    return load_golden
