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


def elem_rel_err(got, ref, frac=1e-2):
    """Per-element relative error max_i |got_i - ref_i| / |ref_i| over the elements with |ref_i| >= frac * max|ref|
    (north_star's "within 1e-4 relative", element by element; scores closer to zero than 1 % of the batch's scale
    are covered by :func:`rel_err`'s scale-normalised bound).  Returns (error, number of elements compared)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if not ref.size:
        return 0.0, 0
    scale = np.abs(ref).max()
    if scale == 0.0:  # an all-zero batch (e.g. every ReLU closed): only exact zeros agree
        return (0.0 if not np.any(got) else float("inf")), int(ref.size)
    keep = np.abs(ref) >= frac * scale
    return float(np.max(np.abs(got - ref)[keep] / np.abs(ref)[keep])), int(keep.sum())


def parity_report(tag, **fields):
    """Append one line to gpurun_out/parity_report.jsonl (copied to profiles/ after a GPU run)."""
    out = REPO / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / "parity_report.jsonl", "a") as fh:
            fh.write(json.dumps({"case": tag, **fields}) + "\n")
    except OSError:
        pass
