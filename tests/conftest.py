import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dnn-inference-engine_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_libs():
    """Builds the checker libraries (a no-op when they are current)."""
    import build as oracle_build
    oracle_build.build_oracle()
    return oracle_build.build_ref()


def normalised_error(out, ref):
    """max|out-ref| / max|ref| -- the parity measure of SURVEY.md section 8(d)."""
    import numpy as np
    return float(np.abs(out.astype(np.float64) - ref.astype(np.float64)).max() / np.abs(ref).max())
