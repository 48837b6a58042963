#!/usr/bin/env python
"""PSK demod timing probe (used under ncu for the per-kernel launch list): python tools/psk_probe.py [log2n] [order] [noise]
noise 0.2 leaves ~3e-4 of the gap samples above the gate (worst case for the speculation: the loop state creeps through the
gaps one spike at a time, serially); 0.3 is what detect_noise_level would pick for this capture (no spikes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C

import numpy as np


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
    from urh_b200 import _lib
    from urh_b200.device import to_device
    from urh_b200.cythonext import signal_functions as sf
    from test_gpu_costas import synth_psk
    ctx = _lib.default_context()
    n = 1 << log2n
    base = synth_psk(min(n, 1 << 22), order, seed=3)
    iq = np.tile(base, (n // len(base), 1))
    d = to_device(iq, ctx)
    for _ in range(2):
        sf.afp_demod(d, noise, "PSK", order)
    ctx.sync()
    ts = []
    for _ in range(3):
        ctx.timer_start()
        sf.afp_demod(d, noise, "PSK", order)
        ts.append(ctx.timer_stop())
    st = (C.c_int64 * 3)()
    ctx.lib.urh_costas_stats(ctx.handle, st)
    print("n=2^%d order=%d noise=%g ms=%s MS/s=%.1f chunks fast/slow/stepped=%s super-chunks redone=%d" % (
        log2n, order, noise, ["%.3f" % t for t in ts], n / np.median(ts) / 1e3, list(st), ctx.lib.urh_costas_last_redone(ctx.handle)))


if __name__ == "__main__":
    main()
