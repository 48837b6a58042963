"""CUDA drop-in for the IQ part of ``urh.cythonext.util`` (reference: src/urh/cythonext/util.pyx)."""
import ctypes as C

import numpy as np

from .. import _lib
from ..device import DeviceArray, to_device


def minmax(arr):
    """util.pyx:20-36 — (min, max) as Python numbers; (0, 0) for an empty array.  Tiny host arrays
    (chunk means, histogram inputs) — plain numpy; device arrays are reduced on the GPU by the callers."""
    if isinstance(arr, DeviceArray):
        arr = arr.get()
    if len(arr) == 0:
        return 0, 0
    return arr.min().item(), arr.max().item()


def get_magnitudes(arr):
    """util.pyx:128-136 — float64[n] magnitudes of an (n,2) IQ array."""
    on_device = isinstance(arr, DeviceArray)
    ctx = arr.ctx if on_device else _lib.default_context()
    if not on_device:
        arr = np.ascontiguousarray(arr)
        if arr.ndim != 2 or arr.dtype not in _lib._DTYPE_CODE:
            raise TypeError("No matching signature found")
    n = len(arr)
    out = DeviceArray(ctx, (n,), np.float64)
    if n:
        d = arr if on_device else to_device(arr, ctx)
        ctx.check(ctx.lib.urh_get_magnitudes(ctx.handle, C.c_void_p(d.ptr), _lib.dtype_code(d.dtype), n, C.c_void_p(out.ptr)))
    return out if on_device else out.get()


def arr2decibel(arr):
    """util.pyx:38-48 — 10*log10(|x|^2) of a 2-D complex64 array, float32."""
    on_device = isinstance(arr, DeviceArray)
    ctx = arr.ctx if on_device else _lib.default_context()
    if not on_device:
        arr = np.ascontiguousarray(arr, dtype=np.complex64)
    shape = arr.shape
    count = int(np.prod(shape))
    out = DeviceArray(ctx, shape, np.float32)
    if count:
        d = arr if on_device else to_device(arr.view(np.float32).reshape(-1), ctx)
        ctx.check(ctx.lib.urh_arr2decibel(ctx.handle, C.c_void_p(d.ptr), count, C.c_void_p(out.ptr)))
    return out if on_device else out.get()


def bit_array_to_number(bits, end: int, start: int = 0) -> int:
    """util.pyx:50-61 (MSB first; host helper used by the modulator set-up)."""
    if end < 1:
        return 0
    r, acc = 0, 1
    for i in range(start, end):
        r += int(bits[end - 1 - i + start]) * acc
        acc *= 2
    return r & 0xFFFFFFFFFFFFFFFF
