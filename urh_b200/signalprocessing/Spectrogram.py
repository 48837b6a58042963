"""``Spectrogram`` numerics (reference: src/urh/signalprocessing/Spectrogram.py:94-206).  STFT + dB on the GPU
(spectrogram.cu: one fused kernel for power-of-two windows, cuFFT for the FFT only otherwise) and the BGRA colormap look-up.
The QImage wrapping of the reference class is GUI and out of scope."""
import ctypes as C
import math

import numpy as np

from .. import _lib
from ..device import DeviceArray, to_device
from .IQArray import IQArray


class Spectrogram(object):
    MAX_LINES_PER_VIEW = 1000
    DEFAULT_FFT_WINDOW_SIZE = 1024

    def __init__(self, samples: np.ndarray, window_size=DEFAULT_FFT_WINDOW_SIZE, overlap_factor=0.5, window_function=np.hanning):
        self.__samples = np.zeros(1, dtype=np.complex64)
        self.samples = samples
        self.window_size = window_size
        self.overlap_factor = overlap_factor
        self.window_function = window_function
        self.data_min, self.data_max = -140, 10

    @property
    def samples(self):
        return self.__samples

    @samples.setter
    def samples(self, value):
        if isinstance(value, IQArray):
            value = value.as_complex64()
        elif isinstance(value, np.ndarray) and value.dtype != np.complex64:
            value = IQArray(value).as_complex64()
        elif value is None:
            value = np.zeros(1, dtype=np.complex64)
        self.__samples = value

    @property
    def time_bins(self):
        return int(math.ceil(len(self.samples) / self.hop_size))

    @property
    def freq_bins(self):
        return self.window_size

    @property
    def hop_size(self):
        return self.window_size - int(self.overlap_factor * self.window_size)

    def _num_frames(self, n):
        return max(1, (max(n, self.window_size) - self.window_size) // self.hop_size + 1)

    def _run(self, samples, mode):
        ctx = _lib.default_context()
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        W, hop = int(self.window_size), int(self.hop_size)
        frames = self._num_frames(len(x))
        window = np.ascontiguousarray(self.window_function(W), dtype=np.float64)
        d_x = to_device(x.view(np.float32) if len(x) else np.zeros(2, np.float32), ctx)
        d_w = to_device(window, ctx)
        if mode == 0:
            out = DeviceArray(ctx, (frames, W), np.complex128)
            ctx.check(ctx.lib.urh_stft(ctx.handle, C.c_void_p(d_x.ptr), len(x), W, hop, C.c_void_p(d_w.ptr), frames, C.c_void_p(out.ptr)))
        else:
            out = DeviceArray(ctx, (frames, W), np.float32)
            ctx.check(ctx.lib.urh_spectrogram_db(ctx.handle, C.c_void_p(d_x.ptr), len(x), W, hop, C.c_void_p(d_w.ptr), frames, C.c_void_p(out.ptr)))
        return out.get()

    def stft(self, samples: np.ndarray):
        """fft(frames * window) / window_size, complex128 [num_frames, window_size] (Spectrogram.py:94-116)"""
        return self._run(samples, 0)

    def calculate_spectrogram(self, samples: np.ndarray = None) -> np.ndarray:
        """fliplr(arr2decibel(fftshift(stft).astype(complex64))), float32 (Spectrogram.py:156-162)"""
        return self._run(self.samples if samples is None else samples, 1)

    @staticmethod
    def apply_bgra_lookup(data: np.ndarray, colormap, data_min=None, data_max=None, normalize=True) -> np.ndarray:
        """Spectrogram.py:192-206 on the GPU: uint8 [cols, rows, 4] image of ``data.T`` through ``colormap`` (entries x 4 bytes BGRA)"""
        if normalize and (data_min is None or data_max is None):
            raise ValueError("Can't normalize without data min and data max")
        ctx = _lib.default_context()
        on_device = isinstance(data, DeviceArray)
        d = data if on_device else to_device(np.ascontiguousarray(data, dtype=np.float32), ctx)
        rows, cols = d.shape
        cmap = np.ascontiguousarray(colormap, dtype=np.uint8)
        if cmap.ndim != 2 or cmap.shape[1] != 4:
            raise ValueError("colormap must be entries x 4 bytes (blue, green, red, alpha)")
        d_map = to_device(cmap, ctx)
        out = DeviceArray(ctx, (cols, rows, 4), np.uint8)
        ctx.check(ctx.lib.urh_bgra_lookup(ctx.handle, C.c_void_p(d.ptr), rows, cols, C.c_void_p(d_map.ptr), len(cmap),
                                          float(data_min) if normalize else 0.0, float(data_max) if normalize else 1.0, int(bool(normalize)),
                                          C.c_void_p(out.ptr)))
        return out if on_device else out.get()
