"""Haar continuous wavelet transform in the frequency domain (reference: src/urh/ainterpretation/Wavelet.py:7-43) on
the GPU: forward FFT (float32 for complex64 input, as numpy >= 2 computes it; double otherwise), multiplication with the
analytic Haar spectrum in complex128, inverse FFT in double (modulation.cu; cuFFT for the FFTs only)."""
import ctypes as C

import numpy as np

from .. import _lib
from ..device import DeviceArray, to_device


def cwt_haar(x, scale=10):
    """-> complex128[P - 4*scale] with P = 2**floor(log2(len(x))) (numpy array)"""
    on_device = isinstance(x, DeviceArray)
    if on_device:
        ctx, d = x.ctx, x
        is_c128 = x.dtype == np.complex128
        if x.dtype not in (np.complex64, np.complex128):
            raise ValueError("cwt_haar on the device takes complex64 or complex128")
    else:
        x = np.asarray(x)
        is_c128 = x.dtype != np.complex64
        x = np.ascontiguousarray(x, dtype=np.complex128 if is_c128 else np.complex64)
        ctx = _lib.default_context()
        d = to_device(x.view(np.float64 if is_c128 else np.float32), ctx) if len(x) else None
    n = len(x)
    if n == 0:
        raise ValueError("cwt_haar of an empty array")  # int(np.log2(0)) raises in the reference as well
    num = 2 ** int(np.log2(n))
    out_len = max(num - 4 * int(scale), 0)
    out = DeviceArray(ctx, (max(out_len, 1),), np.complex128)
    got = C.c_int64(0)
    ctx.check(ctx.lib.urh_cwt_haar(ctx.handle, C.c_void_p(d.ptr), int(is_c128), n, int(scale), C.c_void_p(out.ptr), C.byref(got)))
    return out.get()[: got.value]
