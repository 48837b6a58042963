"""GPU parity: detect_modulation / cwt_haar (modulation.cu, cuFFT for the FFTs) vs golden vectors from the reference and
vs the oracle's numpy restatement.  The forward transform of a complex64 message is a float32 FFT in numpy >= 2 and in
cuFFT C2C alike, but with different rounding, hence tolerances (stated per assertion), decisions compared exactly."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from test_oracle_modulation import golden_modulation, message

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def AI():
    from urh_b200.ainterpretation import AutoInterpretation
    return AutoInterpretation


def test_cwt_haar_golden():
    from urh_b200.ainterpretation import Wavelet
    g, _ = golden_modulation()
    x = g["cwt_x"]
    for key, arg, scale, tol in (("cwt_c64_scale4", x, 4, 2e-5), ("cwt_c128_scale10", x.astype(np.complex128), 10, 1e-11)):
        ref = g[key]
        got = Wavelet.cwt_haar(arg, scale=scale)
        assert got.dtype == np.complex128 and got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= tol * np.max(np.abs(ref)), key


def test_detect_modulation_golden(AI):
    _, index = golden_modulation()
    for rec in index:
        data = message(rec)
        assert AI.detect_modulation(data) == rec["decision"], rec
        feat, spec = AI.modulation_features(data)
        if rec["features"] is not None:
            # float32 FFT rounding: relative 1e-4 on the variances (values span 1e-2 .. 1e1)
            assert np.allclose(feat[3:7], rec["features"], rtol=1e-4, atol=1e-7), (rec, feat)


def test_features_vs_oracle_random(AI, oracle):
    rng = np.random.default_rng(21)
    for trial in range(12):
        n = int(rng.integers(40, 20000))
        t = np.arange(n)
        kind = trial % 4
        if kind == 0:    # FSK-like
            f = np.repeat(rng.choice([-0.05, 0.05], n // 50 + 1), 50)[:n]
            x = np.exp(2j * np.pi * np.cumsum(f))
        elif kind == 1:  # ASK-like
            x = (np.repeat(rng.integers(0, 2, n // 40 + 1), 40)[:n] * 0.8 + 0.2) * np.exp(2j * np.pi * 0.01 * t)
        elif kind == 2:  # PSK-like
            x = np.exp(1j * np.pi * np.repeat(rng.integers(0, 2, n // 30 + 1), 30)[:n]) * np.exp(2j * np.pi * 0.002 * t)
        else:            # a single carrier burst
            x = np.exp(2j * np.pi * 0.03 * t)
        x = (x + 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
        if trial % 5 == 0:
            x[rng.integers(0, n, 2)] = 0   # up to 3 exact zeros are dropped, more mean "OOK"
        nz, ofeat = oracle.modulation_features(x)
        feat, spec = AI.modulation_features(x)
        assert int(feat[0]) == nz
        if ofeat is None:
            assert int(feat[2]) == 0 or len(x) - nz > 3
            continue
        assert np.allclose(feat[3:7], ofeat[:4], rtol=2e-4, atol=1e-7), (trial, feat[3:7], ofeat)
        assert AI._fsk_peak_test(spec) == ofeat[4], trial
        assert AI.detect_modulation(x) == oracle.detect_modulation(x), trial


def test_degenerate_messages(AI, oracle):
    z = np.zeros(50, np.complex64)
    assert AI.detect_modulation(z) is None
    x = np.ones(50, np.complex64)
    x[:10] = 0
    assert AI.detect_modulation(x) == "OOK"                       # more than 3 zeros
    short = (np.arange(1, 18) * (1 + 1j)).astype(np.complex64)     # P = 16 <= 4*scale: no wavelet output
    assert AI.detect_modulation(short) is None and oracle.detect_modulation(short) is None
    one = np.array([1 + 1j], np.complex64)
    assert AI.detect_modulation(one) == oracle.detect_modulation(one)


@pytest.mark.parametrize("name", ["fsk", "ask", "enocean", "homematic"])
def test_estimate_uses_device_modulation(AI, name):
    """the orchestrator end to end: same modulation as the reference's estimate() stored with the golden capture"""
    from urh_b200.signalprocessing.IQArray import IQArray
    g = load_golden("capture_" + name)
    est = AI.estimate(IQArray(g["iq"]))
    ref = g["meta"]["estimate"]
    assert (est is None) == (ref is None)
    if est is not None:
        assert est["modulation_type"] == ref["modulation_type"]
        assert est["bit_length"] == ref["bit_length"]
