import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
MODES = {"inf": 0, "sup": 1, "cen": 2}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(GOLDEN_DIR, "golden.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_x8():
    """8-bit device decimators (oracle/gen_golden_x8.py): (arrays, meta)."""
    z = np.load(os.path.join(GOLDEN_DIR, "golden_x8.npz"))
    with open(os.path.join(GOLDEN_DIR, "golden_x8.json")) as f:
        return {k: z[k] for k in z.files}, json.load(f)


@pytest.fixture(scope="session")
def port():
    """The C restatement oracle (oracle/port), built on demand with gcc."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    from oracle import portbind
    portbind.load()
    return portbind


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; GPU tests fail loudly (never skip to a fallback) if it is missing."""
    from sdrangel_b200 import capi
    capi.lib()
    if capi.device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    capi.init(0)
    return capi


def fnv1a64_u16(a):
    """FNV-1a-64 over uint16 words (SURVEY.md Appendix D convention)."""
    w = np.ascontiguousarray(a).view(np.uint16).ravel()
    h = 1469598103934665603
    # vectorised in chunks is not possible for FNV; use a tight python loop only for small arrays
    for v in w.tolist():
        h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def rel_rms(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    d = np.sqrt(np.mean((a - b) ** 2))
    r = np.sqrt(np.mean(b ** 2))
    return d / r if r > 0 else d


def stream_input():
    """Input of the decim_ii_stream / decim_ii_switch fixtures (oracle/gen_golden.py section 2)."""
    rs = np.random.RandomState(20181018)
    return rs.randint(-32768, 32768, size=2 * 6000).astype(np.int16)
