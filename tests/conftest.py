import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        path = GOLDEN / name
        if name.endswith(".json"):
            return json.load(open(path))
        return np.load(path)

    return load


@pytest.fixture(scope="session")
def native_lib():
    """Build (if needed) and load libflexs_b200.so."""
    from flexs_b200 import _build, _native

    if not _native.lib_path().exists():
        _build.build()
    return _native.lib()


def rel_err(got, ref, floor):
    """max |got-ref| / max(|ref|, floor): the tolerance rule of DESIGN.md (1e-4 relative with an
    absolute floor, because Dense(1) outputs can sit arbitrarily close to zero)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor))) if ref.size else 0.0
