"""GPU: the one-call step (urh_demod_center_digitize: demod -> capture-wide detect_center -> digitize chained on the device,
one host synchronisation) against the oracle and against the call-by-call path whose peak pick runs in numpy."""
import ctypes as C

import numpy as np
import pytest

from conftest import CAPTURES, bits_equal, load_golden, synth_fsk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sf():
    from urh_b200.cythonext import signal_functions

    return signal_functions


def _fsk_wide(n, seed, sps=100, dev=0.05, gaps=True):
    """2-FSK at +-dev cycles/sample (wide enough for a capture-wide center), bursts and gaps"""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, n // sps + 1)
    f = np.repeat(np.where(bits > 0, dev, -dev), sps)[:n]
    x = np.exp(2j * np.pi * np.cumsum(f)) + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    if gaps:
        g = np.arange(n)
        x[(g % 60_000) > 50_000] *= 0.001
        x[int(0.97 * n):] *= 0.001
    return np.ascontiguousarray(np.stack([x.real, x.imag], axis=1).astype(np.float32))


@pytest.mark.parametrize("n", [3, 100, 2047, 2048, 2049, 70_001, 1 << 20, 3_000_001])
@pytest.mark.parametrize("tol", [0, 5, 300])
def test_one_call_matches_oracle(sf, oracle, n, tol):
    iq = _fsk_wide(n, seed=n + tol)
    center, rows, qad = sf.demod_center_digitize(iq, 0.05, "FSK", tol, 100, return_qad=True)
    q_ref = oracle.afp_demod(iq, 0.05, "FSK", 2)
    assert bits_equal(qad, q_ref) == 0
    c_ref = oracle.detect_center(q_ref)
    assert (center is None) == (c_ref is None)
    if center is None:
        assert len(rows) == 0
        return
    assert abs(center - c_ref) <= 2e-6 * max(1.0, abs(c_ref))
    assert np.array_equal(rows, oracle.grab_pulse_lens(q_ref, center, tol, "FSK", 100))
    assert rows[:, 1].sum() == n - tol


@pytest.mark.parametrize("name", CAPTURES)
def test_one_call_golden(sf, oracle, name):
    g = load_golden("capture_" + name)
    mod = g["meta"]["mod"]
    if mod not in ("ASK", "FSK"):
        pytest.skip("ASK/FSK only")
    noise = float(g["noise"])
    center, rows, qad = sf.demod_center_digitize(g["iq"], noise, mod, 5, int(g["meta"].get("sps", 100)), return_qad=True)
    assert bits_equal(qad, g["qad_" + mod]) == 0
    gc = float(g["detect_center"])
    assert center is not None and abs(center - gc) <= 2e-6 * max(1.0, abs(gc))
    sps = int(g["meta"].get("sps", 100))
    assert np.array_equal(rows, oracle.grab_pulse_lens(g["qad_" + mod], center, 5, mod, sps))


@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.int8])
@pytest.mark.parametrize("mod", ["FSK", "ASK"])
def test_one_call_equals_stepwise(sf, dtype, mod):
    n = 1_234_567
    iq = synth_fsk(n, seed=11, gap_every=90_000, dtype=dtype)
    if mod == "ASK":
        env = np.repeat(np.random.default_rng(3).integers(0, 2, n // 100 + 1), 100)[:n] * 0.8 + 0.2
        iq = (iq.astype(np.float32) * env[:, None]).astype(dtype)
    noise = {np.float32: 0.05, np.int16: 1000.0, np.int8: 5.0}[dtype]
    c1, r1, q1 = sf.demod_center_digitize(iq, noise, mod, 5, 100, return_qad=True)
    c2, r2, q2 = sf.demod_center_digitize(iq, noise, mod, 5, 100, return_qad=True, stepwise=True)
    assert bits_equal(q1, q2) == 0
    assert (c1 is None) == (c2 is None)
    if c1 is not None:
        # same histogram; the window's double sums are folded in a different order
        assert abs(c1 - c2) <= 1e-9 * max(1.0, abs(c2))
        assert np.array_equal(r1, sf.grab_pulse_lens(q1, c1, 5, mod, 100))
        if c1 == c2:
            assert np.array_equal(r1, r2)


def test_one_call_constant_and_all_noise(sf):
    n = 50_000
    iq = np.zeros((n, 2), np.float32)   # everything gated: no kept sample -> no center, no rows
    center, rows = sf.demod_center_digitize(iq, 0.05, "FSK", 5, 100)
    assert center is None and len(rows) == 0
    iq[:, 0] = 1.0                      # constant carrier: zero variance -> no center (AutoInterpretation.py:244-247)
    center, rows = sf.demod_center_digitize(iq, 0.05, "FSK", 5, 100)
    assert center is None and len(rows) == 0


def test_one_call_tie_goes_to_numpy(sf):
    """three equally populated peaks: which two are taken is np.argsort's business -> the device hands over to the host path"""
    from urh_b200 import _lib
    from urh_b200.device import DeviceArray, to_device

    ctx = _lib.default_context()
    n = 9 * 4096
    ang = np.repeat(np.array([-1.0, 0.0, 1.0] * 3, dtype=np.float64), 4096)   # three plateaus of equal population
    x = np.exp(1j * np.cumsum(ang))
    iq = np.ascontiguousarray(np.stack([x.real, x.imag], axis=1).astype(np.float32))
    d_iq = to_device(iq, ctx)
    qad = DeviceArray(ctx, (n,), np.float32)
    center, state, k = C.c_double(0.0), C.c_int(0), C.c_int64(0)
    ctx.check(ctx.lib.urh_demod_center_digitize(ctx.handle, C.c_void_p(d_iq.ptr), _lib.DT_F32, n, 0.05, _lib.MOD_FSK, 5, 100, -1,
                                                C.c_void_p(qad.ptr), C.byref(center), C.byref(state), C.byref(k)))
    c1, r1 = sf.demod_center_digitize(iq, 0.05, "FSK", 5, 100)
    c2, r2 = sf.demod_center_digitize(iq, 0.05, "FSK", 5, 100, stepwise=True)
    assert c1 == c2 and np.array_equal(r1, r2)
    assert state.value in (1, 2)   # float rounding may break the tie between the plateaus; if not, the device must not decide
    if state.value == 1:
        assert center.value == c2


def test_launch_count_of_the_step(sf):
    """the whole step is a dozen launches and one synchronisation"""
    from urh_b200 import _lib

    ctx = _lib.default_context()
    iq = _fsk_wide(1 << 20, seed=1)
    sf.demod_center_digitize(iq, 0.05, "FSK", 5, 100)
    before = ctx.launch_count()
    sf.demod_center_digitize(iq, 0.05, "FSK", 5, 100)
    assert ctx.launch_count() - before <= 20


@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.int8])
@pytest.mark.parametrize("chunk", [2048, 100_000, 1 << 24])
def test_streamed_host_input_equals_resident(sf, dtype, chunk):
    """urh_demod_center_digitize_host (chunked upload on the copy stream, each chunk demodulated as it lands) == the resident call"""
    from urh_b200 import _lib
    from urh_b200.device import PinnedArray, to_device

    ctx = _lib.default_context()
    n = 1_000_003
    iq = synth_fsk(n, seed=5, gap_every=70_000, dtype=dtype)
    noise = {np.float32: 0.05, np.int16: 1000.0, np.int8: 5.0}[dtype]
    pinned = PinnedArray(iq.shape, iq.dtype, ctx)
    pinned.array[...] = iq
    rows_buf = PinnedArray((n // 8 + 1024, 2), np.int64, ctx)
    c_host, r_host, q_host = sf.demod_center_digitize(pinned.array, noise, "FSK", 5, 100, return_qad=True, chunk_samples=chunk,
                                                      rows_out=rows_buf.array)
    c_dev, r_dev, q_dev = sf.demod_center_digitize(to_device(iq, ctx), noise, "FSK", 5, 100, return_qad=True)
    assert bits_equal(q_host, q_dev.get()) == 0
    assert c_host == c_dev
    assert np.array_equal(r_host, r_dev)
    # (when the device decides the center, r_host is a view of the caller's pinned buffer; a tie between histogram peaks hands
    # the decision to the host path, which allocates its own table)
    r_host = None
    pinned.free()
    rows_buf.free()
