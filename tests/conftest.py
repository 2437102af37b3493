import os
import sys

import pytest

# route even toy-sized products through the tcgen05 engine so that the parity tests cover its edge cases
os.environ.setdefault("RAU_TC_MIN_WORK", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by `pytest -m gpu` on the GPU box)")


def small_cfg(**kw):
    """A toy configuration the float64 oracle evaluates in milliseconds."""
    from oracle import rau_oracle as O
    base = dict(V=50, embed=12, Hq=16, nlayer=2, C=24, S=20, M=16, A=8, H=16, N=30, nHop=2, T=5)
    base.update(kw)
    return O.RauConfig(**base)
