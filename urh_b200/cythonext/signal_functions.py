"""CUDA drop-in for ``urh.cythonext.signal_functions`` (reference: src/urh/cythonext/signal_functions.pyx)."""
import ctypes as C

import numpy as np

from .. import _lib
from ..device import DeviceArray, to_device

_IQ_DTYPES = (np.int8, np.uint8, np.int16, np.uint16, np.float32)


def _check_iq(samples):
    """Mimic the Cython fused-type dispatch of ``IQ samples`` (util.pxd:1-10)."""
    if isinstance(samples, DeviceArray):
        if samples.ndim != 2 or samples.shape[1] != 2:
            raise ValueError("Buffer has wrong number of dimensions (expected 2)")
        if samples.dtype not in [np.dtype(t) for t in _IQ_DTYPES]:
            raise TypeError("No matching signature found")
        return samples
    samples = np.asarray(samples) if not isinstance(samples, np.ndarray) else samples
    if samples.ndim != 2:
        raise TypeError("No matching signature found")
    if samples.dtype not in [np.dtype(t) for t in _IQ_DTYPES]:
        raise TypeError("No matching signature found")
    if samples.shape[1] != 2 and len(samples):
        raise ValueError("IQ samples must have shape (n, 2)")
    if not samples.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return samples


def afp_demod(samples, noise_mag: float, mod_type: str, mod_order: int, costas_loop_bandwidth: float = 0.1):
    """signal_functions.pyx:333-378.  Returns float32[n] (numpy for numpy input, DeviceArray for device input)."""
    samples = _check_iq(samples)
    on_device = isinstance(samples, DeviceArray)
    ctx = samples.ctx if on_device else _lib.default_context()
    n = len(samples)
    if n == 0:
        return DeviceArray(ctx, (0,), np.float32) if on_device else np.zeros(0, dtype=np.float32)
    d_iq = samples if on_device else to_device(samples, ctx)
    out = DeviceArray(ctx, (n,), np.float32)
    code = _lib.demod_mod_code(mod_type)
    ctx.check(
        ctx.lib.urh_afp_demod(
            ctx.handle, C.c_void_p(d_iq.ptr), _lib.dtype_code(d_iq.dtype), n, float(noise_mag), code if code >= 0 else 99,
            int(mod_order), float(costas_loop_bandwidth), C.c_void_p(out.ptr),
        )
    )
    return out if on_device else out.get()


def get_center_thresholds(center: float, spacing: float, modulation_order: int) -> np.ndarray:
    """signal_functions.pyx:380-390 (host arithmetic; float32)."""
    lib = _lib.load_library()
    out = np.empty(max(int(modulation_order) - 1, 0), dtype=np.float32)
    if len(out):
        lib.urh_get_center_thresholds(float(center), float(spacing), int(modulation_order), out.ctypes.data_as(C.c_void_p))
    return out


def _fetch_pulses(ctx, k: int) -> np.ndarray:
    rows = np.empty((k, 2), dtype=np.int64)
    if k:
        ctx.check(ctx.lib.urh_fetch_pulses(ctx.handle, rows.ctypes.data_as(C.c_void_p), k))
    return rows


def grab_pulse_lens(samples, center: float, tolerance: int, modulation_type: str, samples_per_symbol: int,
                    bits_per_symbol: int = 1, center_spacing: float = 0.1) -> np.ndarray:
    """signal_functions.pyx:392-495.  ``samples`` float32[n] (numpy or DeviceArray) -> int64[k,2] (numpy)."""
    on_device = isinstance(samples, DeviceArray)
    if not on_device:
        if not isinstance(samples, np.ndarray) or samples.dtype != np.float32 or samples.ndim != 1:
            raise ValueError("Buffer dtype mismatch, expected 'float' (1-D float32)")
        if not samples.flags.c_contiguous:
            raise ValueError("ndarray is not C-contiguous")
    elif samples.dtype != np.float32 or samples.ndim != 1:
        raise ValueError("Buffer dtype mismatch, expected 'float' (1-D float32)")
    if not 0 <= int(tolerance) <= 0xFFFF:
        raise OverflowError("value too large to convert to uint16_t")
    if not 0 <= int(bits_per_symbol) <= 0xFF:
        raise OverflowError("value too large to convert to uint8_t")
    ctx = samples.ctx if on_device else _lib.default_context()
    n = len(samples)
    if n == 0:
        return np.zeros((0, 2), dtype=np.int64)
    d = samples if on_device else to_device(samples, ctx)
    k = C.c_int64(0)
    code = _lib.demod_mod_code(modulation_type)
    ctx.check(
        ctx.lib.urh_grab_pulse_lens(
            ctx.handle, C.c_void_p(d.ptr), n, float(center), int(tolerance), code if code >= 0 else 99,
            int(samples_per_symbol), int(bits_per_symbol), float(center_spacing), C.byref(k),
        )
    )
    return _fetch_pulses(ctx, k.value)


def demod_digitize(samples, noise_mag: float, mod_type: str, center: float, tolerance: int, samples_per_symbol: int,
                   bits_per_symbol: int = 1, center_spacing: float = 0.1, return_qad: bool = True):
    """Fused afp_demod + grab_pulse_lens in ONE pass over the IQ samples (B200 addition; same results as
    calling the two reference functions back to back).  Returns (qad or None, int64[k,2])."""
    samples = _check_iq(samples)
    on_device = isinstance(samples, DeviceArray)
    ctx = samples.ctx if on_device else _lib.default_context()
    n = len(samples)
    if n == 0:
        return (np.zeros(0, np.float32) if return_qad else None), np.zeros((0, 2), dtype=np.int64)
    d_iq = samples if on_device else to_device(samples, ctx)
    qad = DeviceArray(ctx, (n,), np.float32) if return_qad else None
    k = C.c_int64(0)
    code = _lib.demod_mod_code(mod_type)
    ctx.check(
        ctx.lib.urh_demod_digitize(
            ctx.handle, C.c_void_p(d_iq.ptr), _lib.dtype_code(d_iq.dtype), n, float(noise_mag), code if code >= 0 else 99,
            float(center), int(tolerance), int(samples_per_symbol), int(bits_per_symbol), float(center_spacing),
            C.c_void_p(qad.ptr if qad is not None else 0), C.byref(k),
        )
    )
    rows = _fetch_pulses(ctx, k.value)
    if qad is not None and not on_device:
        qad = qad.get()
    return qad, rows
