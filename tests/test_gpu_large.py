"""GPU: index arithmetic past 2^31 BYTES of IQ.  A capture of 2^28 + 12345 complex64 samples (2.1 GB; qad positions, staging
slots, tile * stage_cap products and byte offsets all cross 32-bit boundaries) through the one-call step and the fused
known-center path, checked against the CPU oracle on windows that lie beyond the 2^31-byte mark and through size-independent
properties of the whole table."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = (1 << 28) + 12345
SPS, TOL, NOISE = 100, 5, 0.05


@pytest.fixture(scope="module")
def capture():
    from urh_b200 import _lib
    from urh_b200.device import DeviceArray

    ctx = _lib.default_context()
    rng = np.random.Generator(np.random.PCG64(7))
    nsym = N // SPS + 2
    b = (rng.integers(0, 2, nsym, dtype=np.int8) * 2 - 1).astype(np.int8)
    s = np.zeros(nsym, dtype=np.int32)
    np.cumsum(b[:-1], out=s[1:], dtype=np.int32)
    d_b = DeviceArray(ctx, (nsym,), np.int8).set(b)
    d_s = DeviceArray(ctx, (nsym,), np.int32).set(s)
    d_iq = DeviceArray(ctx, (N, 2), np.float32)
    ctx.check(ctx.lib.urh_synth_fsk(ctx.handle, C.c_void_p(d_iq.ptr), N, 0, SPS, C.c_void_p(d_b.ptr), C.c_void_p(d_s.ptr),
                                    C.c_double(0.05), 1.0, 0.01, 4711, 6_000_000, 5_000_000, int(0.40 * N), int(0.43 * N), int(0.97 * N)))
    ctx.sync()
    yield ctx, d_iq
    d_iq.free()


def _boundaries(rows, first_offset):
    """(position of the firing that ends each row, state) for every row but the tail"""
    pos = first_offset + TOL - 1 + np.cumsum(rows[:-1, 1])
    return pos, rows[:-1, 0]


def test_one_call_step_beyond_2_31_bytes(capture, oracle):
    from urh_b200.cythonext import signal_functions as sf
    from urh_b200.device import DeviceArray

    ctx, d_iq = capture
    qad = DeviceArray(ctx, (N,), np.float32)
    center, rows = sf.demod_center_digitize(d_iq, NOISE, "FSK", TOL, SPS, out=qad)
    assert center is not None
    assert int(rows[:, 1].sum()) == N - TOL                       # the pulse lengths tile the capture
    assert np.all(rows[1:, 0] != rows[:-1, 0])                    # neighbours differ (FSK: no relabelling)
    pos_gpu, st_gpu = _boundaries(rows, 0)
    w = 1 << 22
    for a in ((1 << 28) - w // 2, (1 << 28) + 12345 - w, (1 << 27) + 999):   # across the 2^31-byte mark, the tail, and 2^30 bytes
        iq = d_iq[a - 1: a + w].get()
        q_ref = oracle.afp_demod(np.ascontiguousarray(iq), NOISE, "FSK", 2)[1:]
        q_gpu = qad[a: a + w].get()
        assert int(np.count_nonzero(q_gpu.view(np.uint32) != q_ref.view(np.uint32))) == 0
        r_ref = oracle.grab_pulse_lens(q_ref, center, TOL, "FSK", SPS)
        p_ref, s_ref = _boundaries(r_ref, a)
        lo, hi = a + (1 << 20), a + w
        mg, mr = (pos_gpu > lo) & (pos_gpu < hi), (p_ref > lo) & (p_ref < hi)
        assert np.array_equal(pos_gpu[mg], p_ref[mr]) and np.array_equal(st_gpu[mg], s_ref[mr])
    # the fused known-center path on the same capture: same table as the two-step result for that center
    q2, rows_fused = sf.demod_digitize(d_iq, NOISE, "FSK", float(center), TOL, SPS, return_qad=False)
    assert np.array_equal(rows_fused, rows)
    qad.free()


def test_stepwise_and_bitwise_center_at_size(capture, oracle):
    """detect_center on 2^28 demodulated samples: the stand-alone path (numpy's float32 pairwise variance replayed: compaction and
    tree over 2.4e8 window samples) returns the oracle's center bit for bit"""
    from urh_b200.ainterpretation import AutoInterpretation as AI
    from urh_b200.cythonext import signal_functions as sf

    ctx, d_iq = capture
    qad = sf.afp_demod(d_iq, NOISE, "FSK", 2)
    mine = AI.detect_center(qad)
    ref = oracle.detect_center(qad.get())
    assert mine is not None and float(mine) == float(ref)
    qad.free()
