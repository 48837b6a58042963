"""``Filter`` (reference: src/urh/signalprocessing/Filter.py).  Same class / static methods; the convolutions and
the DC correction run on the GPU (filter.cu)."""
import ctypes as C
import math
from enum import Enum

import numpy as np

from .. import _lib, settings
from ..cythonext import signal_functions
from ..device import DeviceArray, to_device


class FilterType(Enum):
    moving_average = "moving average"
    dc_correction = "DC correction"
    custom = "custom"


class Filter(object):
    BANDWIDTHS = {"Very Narrow": 0.001, "Narrow": 0.01, "Medium": 0.08, "Wide": 0.1, "Very Wide": 0.42}
    # up to this many rows the DC correction reproduces numpy's serial float32 column sums bit for bit; beyond it an
    # accurate double reduction is used (the reference's own mean is off by percent there)
    EXACT_DC_MAX = 1 << 22

    def __init__(self, taps: list, filter_type: FilterType = FilterType.custom):
        self.filter_type = filter_type
        self.taps = taps

    def work(self, input_signal: np.ndarray) -> np.ndarray:
        if self.filter_type == FilterType.dc_correction:
            return self.dc_correction(input_signal)
        return self.apply_fir_filter(input_signal.flatten())

    @staticmethod
    def dc_correction(input_signal: np.ndarray) -> np.ndarray:
        """input_signal - np.mean(input_signal, axis=0) (Filter.py:32-33)"""
        ctx = _lib.default_context()
        if input_signal.dtype != np.float32:
            # integer captures: numpy promotes to float64 and the column means of integers are exact in double
            x = np.ascontiguousarray(input_signal)
            if x.dtype not in (np.int8, np.uint8, np.int16, np.uint16) or x.ndim != 2 or x.shape[1] != 2:
                raise ValueError("dc_correction expects an (n, 2) capture of int8/uint8/int16/uint16/float32")
            if len(x) == 0:
                return x.astype(np.float64)
            d = to_device(x, ctx)
            out = DeviceArray(ctx, x.shape, np.float64)
            ctx.check(ctx.lib.urh_dc_correction_int(ctx.handle, C.c_void_p(d.ptr), _lib.dtype_code(x.dtype), len(x), C.c_void_p(out.ptr)))
            return out.get()
        ctx = _lib.default_context()
        x = np.ascontiguousarray(input_signal)
        n = len(x)
        if n == 0:
            return x.copy()
        d = to_device(x, ctx)
        out = DeviceArray(ctx, x.shape, np.float32)
        ctx.check(ctx.lib.urh_dc_correction(ctx.handle, C.c_void_p(d.ptr), n, C.c_void_p(out.ptr), int(n <= Filter.EXACT_DC_MAX)))
        return out.get()

    def apply_fir_filter(self, input_signal: np.ndarray) -> np.ndarray:
        if input_signal.dtype != np.complex64:
            tmp = np.empty(len(input_signal) // 2, dtype=np.complex64)
            tmp.real = input_signal[0::2]
            tmp.imag = input_signal[1::2]
            input_signal = tmp
        return signal_functions.fir_filter(input_signal, np.array(self.taps, dtype=np.complex64))

    @staticmethod
    def read_configured_filter_bw() -> float:
        bw_type = settings.read("bandpass_filter_bw_type", "Medium", str)
        if bw_type in Filter.BANDWIDTHS:
            return Filter.BANDWIDTHS[bw_type]
        if bw_type.lower() == "custom":
            return settings.read("bandpass_filter_custom_bw", 0.1, float)
        return 0.08

    @staticmethod
    def get_bandwidth_from_filter_length(N):
        return 4 / N

    @staticmethod
    def get_filter_length_from_bandwidth(bw):
        N = int(math.ceil((4 / bw)))
        return N + 1 if N % 2 == 0 else N  # odd length

    @staticmethod
    def _convolve_full_slice(data: np.ndarray, h: np.ndarray, offset: int, out_len: int) -> np.ndarray:
        """full_convolution(data, h)[offset : offset + out_len] on the GPU (complex128 taps, double accumulation)"""
        ctx = _lib.default_context()
        x = np.ascontiguousarray(data, dtype=np.complex64)
        taps = np.ascontiguousarray(h, dtype=np.complex128)
        d_x = to_device(x.view(np.float32), ctx)
        d_t = to_device(taps.view(np.float64), ctx)
        out = DeviceArray(ctx, (out_len,), np.complex64)
        ctx.check(ctx.lib.urh_convolve_c128(ctx.handle, C.c_void_p(d_x.ptr), len(x), C.c_void_p(d_t.ptr), len(taps), int(offset),
                                            int(out_len), C.c_void_p(out.ptr)))
        return out.get()

    @staticmethod
    def fft_convolve_1d(x: np.ndarray, h: np.ndarray):
        """Filter.py:69-82 — centred crop of the full convolution (the reference computes it with a power-of-two FFT in
        complex128; here it is a direct convolution on the GPU with complex128 taps and double accumulation, returned as
        complex64 — what every reference caller casts the result to (IQArray) — i.e. the values agree to 1e-5 of the signal
        scale, DESIGN.md 4.5, not to float64 precision).  len(h) <= 2 gives too_much == 0 and the reference's
        ``result[0:-0]`` is EMPTY: reproduced."""
        n = len(x) + len(h) - 1
        too_much = (n - len(x)) // 2
        if too_much == 0:
            return np.zeros(0, dtype=np.complex64)
        return Filter._convolve_full_slice(x, h, too_much, n - 2 * too_much)

    @staticmethod
    def apply_bandpass_filter(data, f_low, f_high, filter_bw=0.08):
        if f_low > f_high:
            f_low, f_high = f_high, f_low
        f_low = max(-0.5, min(0.5, f_low))
        f_high = max(-0.5, min(0.5, f_high))
        h = Filter.design_windowed_sinc_bandpass(f_low, f_high, filter_bw)
        if len(h) < 8 * math.log(math.sqrt(len(data))):
            # np.convolve(data, h, "same"): centred on the longer operand
            big, small = max(len(data), len(h)), min(len(data), len(h))
            return Filter._convolve_full_slice(data, h, (small - 1) // 2, big)
        return Filter.fft_convolve_1d(data, h)

    @staticmethod
    def design_windowed_sinc_lpf(fc, bw):
        N = Filter.get_filter_length_from_bandwidth(bw)
        h = np.sinc(2 * fc * (np.arange(N) - (N - 1) / 2.0)) * np.blackman(N)
        return h / np.sum(h)  # unity gain

    @staticmethod
    def design_windowed_sinc_bandpass(f_low, f_high, bw):
        f_shift = (f_low + f_high) / 2
        f_c = (f_high - f_low) / 2
        N = Filter.get_filter_length_from_bandwidth(bw)
        return Filter.design_windowed_sinc_lpf(f_c, bw=bw) * np.exp(complex(0, 1) * np.pi * 2 * f_shift * np.arange(0, N, dtype=complex))
