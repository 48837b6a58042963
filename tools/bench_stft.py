#!/usr/bin/env python
"""Spectrogram STFT -> dB (W = 1024, hop 512): the fused shared-memory-FFT kernel against the cuFFT path (URH_B200_STFT_CUFFT=1),
and their agreement.   python tools/bench_stft.py [--log2n 28]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from urh_b200 import _lib
    from urh_b200.device import DeviceArray, to_device

    ctx = _lib.default_context()
    lib = ctx.lib
    n = 1 << args.log2n
    rng = np.random.default_rng(0)
    chunk = (rng.standard_normal((1 << 20, 2)) * 0.1).astype(np.float32)
    chunk[:, 0] += np.cos(2 * np.pi * 0.05 * np.arange(1 << 20)).astype(np.float32)
    chunk[:, 1] += np.sin(2 * np.pi * 0.05 * np.arange(1 << 20)).astype(np.float32)
    d_x = DeviceArray(ctx, (n, 2), np.float32)
    d_c = to_device(chunk, ctx)
    for i in range(n >> 20):
        ctx.check(lib.urh_memcpy_d2d(ctx.handle, C.c_void_p(d_x.ptr + i * chunk.nbytes), C.c_void_p(d_c.ptr), chunk.nbytes))
    W, hop = 1024, 512
    frames = (n - W) // hop + 1
    d_w = to_device(np.hanning(W), ctx)
    out = {}
    res = {}
    for name, env in (("fused", None), ("cufft", "1")):
        if env:
            os.environ["URH_B200_STFT_CUFFT"] = env
        else:
            os.environ.pop("URH_B200_STFT_CUFFT", None)
        d_db = DeviceArray(ctx, (frames, W), np.float32)
        ms = []
        for rep in range(args.reps + 1):
            ctx.timer_start()
            ctx.check(lib.urh_spectrogram_db(ctx.handle, C.c_void_p(d_x.ptr), n, W, hop, C.c_void_p(d_w.ptr), frames, C.c_void_p(d_db.ptr)))
            t = ctx.timer_stop()
            if rep:
                ms.append(t)
        out[name + "_ms"] = float(np.median(ms))
        res[name] = d_db[:4096].get()
        d_db.free()
    peak = res["cufft"].max()
    mask = res["cufft"] > peak - 100
    out["max_abs_dB_diff_within_100dB_of_peak"] = float(np.abs(res["fused"] - res["cufft"])[mask].max())
    out["samples"] = n
    out["algorithmic_GBps_fused"] = 16.0 * n / out["fused_ms"] / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
