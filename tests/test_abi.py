"""The C-ABI library loads on a CPU-only box and exports exactly what include/sia_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "sia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sia_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(native_lib):
    from shazam_b200 import _native
    names = header_functions()
    assert len(names) >= 20
    raw = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in sia_b200.h but not exported"
        assert n in _native.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_native.SIGNATURES) == names


def test_constants_and_defaults(native_lib):
    from shazam_b200 import _native as N
    assert native_lib.sia_version() >= 100
    p = N.default_params()
    # __init__.py:40-51
    assert (p.Fs, p.wsize, p.wratio, p.fan_value, p.amp_min, p.connectivity, p.nbhd) == (44100.0, 4096, 0.5, 5, 10.0, 2, 10)
    for n in (0, 1, 3000, 4096, 6143, 6144, 8191, 220500, 7938000):
        assert native_lib.sia_num_frames(n) == N.num_frames(n)
    assert N.num_frames(220500) == 106 and N.num_frames(7938000) == 3874


def test_errors_do_not_throw(native_lib):
    from shazam_b200 import _native as N
    # invalid arguments come back as codes + message, no GPU needed
    assert native_lib.sia_ctx_create(0, 0, None) == N.E_INVALID
    assert b"NULL" in native_lib.sia_last_error()
    assert native_lib.sia_ctx_destroy(None) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from shazam_b200.fingerprinter import Fingerprinter
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Fingerprinter(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "shazam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("sia_oracle", "oracle") \
                    or f == "__never__", f"{f} mentions the oracle"
