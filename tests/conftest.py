"""Shared pytest configuration: the `gpu` marker, paths and golden-vector loading."""
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
    # a fresh checkout has no liblbm_b200.so (built artefacts are git-ignored): build it once, with nvcc
    # cross-compiling for sm_100a; an existing library is used as is (the GPU box receives the prebuilt one)
    from lettuce_b200 import build as _build
    if not os.path.exists(_build.LIB):
        _build.build()


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def max_rel(a, b):
    """max |a-b| / |b| the way SURVEY.md section 8c defines the parity metric."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / denom))


@pytest.fixture(scope="session")
def golden():
    return load_golden
