import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def rel_l2(a, b):
    import numpy as np
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = np.asarray(a, dtype=dt)
    b = np.asarray(b, dtype=dt)
    d = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (d if d > 0 else 1.0))
