"""Haar continuous wavelet transform in the frequency domain (reference: src/urh/ainterpretation/Wavelet.py:7-43).

Second-tier row of the scope table: per-message FFTs of at most a few million points.  Round 1
keeps numpy's FFT on the host so that `detect_modulation` has exactly the reference's numerics; the cuFFT
version is the planned replacement.
"""
import numpy as np


def normalized_haar_wavelet(omega, scale):
    scaled = omega[:] / scale
    scaled[0] = 1.0  # omega[0] == 0: avoid 0/0, the numerator is 0 there anyway
    return (1j * np.square(-1 + np.exp(0.5j * omega))) / scaled


def cwt_haar(x: np.ndarray, scale=10):
    num = 2 ** int(np.log2(len(x)))  # truncate to a power of two
    x = x[0:num]
    x_hat = np.fft.fft(x)
    f = 2.0 * np.pi / num
    omega = f * np.concatenate((np.arange(0, num // 2), np.arange(num // 2, num) * -1))
    psi_hat = np.sqrt(2.0 * np.pi * scale) * normalized_haar_wavelet(scale * omega, scale)
    W = np.fft.ifft(x_hat * psi_hat)
    return W[2 * scale: -2 * scale]
