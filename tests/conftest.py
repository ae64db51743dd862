import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU checkers (test infrastructure)."""
    import oracle as O
    O.load_oracle()
    return O


@pytest.fixture(scope="session")
def saf():
    """The product binding; GPU tests fail loudly if the CUDA library is missing."""
    import spatial_audio_framework_b200 as s
    s.lib()
    return s


def golden_files(kind=None):
    out = []
    for p in sorted(GOLDEN.glob("*.npz")):
        k = str(np.load(p)["kind"])
        if k.startswith("producers"):        # fixtures of the filter producers: tests/test_producers_cpu.py, test_gpu_producers.py
            continue
        if kind is None or k == kind:
            out.append(p)
    return out


def err_metrics(y, ref):
    """(max abs error / full scale, relative L2) with full scale = max|ref|."""
    y = np.asarray(y, np.float64)
    ref = np.asarray(ref, np.float64)
    fs = float(np.abs(ref).max())
    d = y - ref
    return float(np.abs(d).max() / fs), float(np.linalg.norm(d) / np.linalg.norm(ref))


# north_star tolerance: max abs error <= 1e-5 of full scale and relative L2 <= 1e-6
TOL_MAXABS_FS = 1e-5
TOL_REL_L2 = 1e-6
