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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def native_lib():
    """The in-tree shared library (built on demand: nvcc cross-compiles without a GPU)."""
    from shazam_b200 import _native, build
    if not os.path.exists(_native.LIB_PATH):
        build.build_native()
    return _native.lib()


@pytest.fixture(scope="session")
def wav_fixture():
    return np.load(os.path.join(GOLDEN, "wav_fixture.npz"))


@pytest.fixture(scope="session")
def synth_cases():
    return np.load(os.path.join(GOLDEN, "synth_cases.npz"))


@pytest.fixture(scope="session")
def peaks_cases():
    return np.load(os.path.join(GOLDEN, "peaks_cases.npz"))


@pytest.fixture(scope="session")
def match_cases():
    return json.load(open(os.path.join(GOLDEN, "match_cases.json")))


@pytest.fixture(scope="session")
def fpr(native_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from shazam_b200.fingerprinter import Fingerprinter
    f = Fingerprinter(0, max_chunk_frames=16384)
    yield f
    f.close()
