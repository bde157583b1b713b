import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def G():
    return golden


def pdf_tolerance(edges, cdf, inds, ulps=6.0):
    """Per-sample tolerance for inverse-CDF outputs computed from a CDF that differs from the
    reference's by summation-order ulps: t=(u-cdf_lo)/denom moves by ~ulps*6e-8/denom, scaled by the
    bin width (sampling_utils.py:61-64)."""
    M1 = cdf.shape[-1]
    below = np.clip(inds - 1, 0, M1 - 1); above = np.clip(inds, 1, M1 - 1)
    denom = np.take_along_axis(cdf, above, -1) - np.take_along_axis(cdf, below, -1)
    denom = np.where(denom < 1e-5, 1.0, denom)
    width = np.abs(np.take_along_axis(edges, above, -1) - np.take_along_axis(edges, below, -1))
    return 2e-6 + width * ulps * 6e-8 / denom
