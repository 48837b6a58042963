"""GPU: the speculative chunk-parallel Costas loop (costas_spec.cu) is bit-identical to the serial recurrence
(oracle == the reference's costa_demod) on long PSK captures with bursts, gaps and noise."""
import ctypes as C
import time

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def synth_psk(n, order, sps=300, sigma=0.05, seed=0, gap_period=600000, gap_len=100000, dtype=np.float32):
    rng = np.random.default_rng(seed)
    nsym = n // sps + 1
    sym = rng.integers(0, order, nsym)
    ph = 2 * np.pi * 0.025 * np.arange(n) + np.repeat(2 * np.pi * sym / order + (np.pi / 4 if order == 4 else 0), sps)[:n]
    on = (np.arange(n) % gap_period) < (gap_period - gap_len)
    x = on * np.exp(1j * ph) + sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty((n, 2), np.float32)
    iq[:, 0] = x.real
    iq[:, 1] = x.imag
    if dtype == np.int8:
        return np.clip(iq * 100, -128, 127).astype(np.int8)
    return iq


@pytest.mark.parametrize("order,dtype,noise", [(2, np.float32, 0.2), (4, np.float32, 0.2), (2, np.int8, 20.0), (2, np.float32, 0.0)])
def test_speculative_costas_bit_exact(oracle, ctx, order, dtype, noise):
    from urh_b200.cythonext import signal_functions as sf

    n = 1_500_000
    iq = synth_psk(n, order, seed=order, dtype=dtype)
    t0 = time.time()
    q = sf.afp_demod(iq, noise, "PSK", order)
    t_gpu = time.time() - t0
    ref = oracle.afp_demod(iq, noise, "PSK", order)
    assert bits_equal(q[1:], ref[1:]) == 0
    st = (C.c_int64 * 3)()
    ctx.lib.urh_costas_stats(ctx.handle, st)
    # the speculation must carry almost all chunks (otherwise the kernel silently degrades to the serial loop)
    assert st[0] >= 0.7 * (n // 4096), list(st)
    print("costas order", order, "gpu %.3fs" % t_gpu, "fast/slow/stepped", list(st))


def test_speculative_costas_edge_sizes(oracle):
    from urh_b200.cythonext import signal_functions as sf

    for n in (4 * 4096, 4 * 4096 + 1, 5 * 4096 - 1, 100_003):
        iq = synth_psk(n, 2, seed=n, gap_period=30000, gap_len=9000)
        assert bits_equal(sf.afp_demod(iq, 0.2, "PSK", 2)[1:], oracle.afp_demod(iq, 0.2, "PSK", 2)[1:]) == 0, n
        iq[:] = 0
        assert bits_equal(sf.afp_demod(iq, 0.2, "PSK", 2)[1:], oracle.afp_demod(iq, 0.2, "PSK", 2)[1:]) == 0
