"""GPU: the packed reciprocal/Newton/residual division of the FSK fast path (fsk_fast.cuh: urh_div2_window)
returns exactly __fdiv_rn's correctly rounded quotient for operands inside its exponent window."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_packed_division_matches_fdiv_rn(ctx):
    for seed in (1, 2, 3):
        bad, tested = C.c_int64(-1), C.c_int64(0)
        ctx.check(ctx.lib.urh_selftest_packed_div(ctx.handle, seed, 1 << 29, C.byref(bad), C.byref(tested)))
        assert tested.value == 2 << 29
        assert bad.value == 0, (seed, bad.value)
