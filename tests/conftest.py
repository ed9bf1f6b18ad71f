import json
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


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "reference_vectors.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def oracle():
    from oracle import fhmc_oracle
    fhmc_oracle.lib()
    return fhmc_oracle


def sel_rows(n, mom):
    """The three quantities the batched API averages by default: N, N^2, U."""
    i = np.arange(n, dtype=np.float64)
    return np.stack([i, i * i, mom[0, 0, 0, 0, 1]])
