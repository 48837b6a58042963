"""CPU suite: the C-ABI library builds, loads and exports exactly what include/urh_b200.h declares,
and the product refuses to run (loudly) without a CUDA device — there is no CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "urh_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(urh_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from urh_b200 import build, _lib

    path = build.build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    # the ctypes table covers the whole header (and nothing that is not declared)
    assert sorted(_lib.SIGNATURES) == syms


def test_no_cpu_fallback_without_gpu():
    from urh_b200 import _lib

    if _lib.cuda_available():
        pytest.skip("GPU present")
    import numpy as np
    from urh_b200.cythonext import signal_functions as sf

    with pytest.raises(_lib.UrhCudaUnavailable):
        sf.afp_demod(np.zeros((10, 2), np.float32), 0.0, "FSK", 2)


def test_product_does_not_import_oracle():
    """No file under urh_b200/ may reference the oracle package."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "urh_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liburh_oracle|oracle/_ref", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_get_center_thresholds_host():
    import numpy as np
    from urh_b200.cythonext import signal_functions as sf
    from oracle import oracle

    for center, spacing, order in [(0.0, 0.1, 2), (0.02, 0.1, 4), (-0.3, 1.5, 8), (0.1234567, 0.333, 16)]:
        assert np.array_equal(sf.get_center_thresholds(center, spacing, order), oracle.get_center_thresholds(center, spacing, order))
