"""CPU suite: the float32 restatement of glibc atan2f used by the FSK kernel (urh_b200/csrc/fdlibm_atan2f.h)
is bit-identical to this box's libm on a broad input sweep (host compilation of the same header)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include "%s"
void restated_atan2f(const float* y, const float* x, float* out, long n) {
    for (long i = 0; i < n; i++) out[i] = urh_atan2f(y[i], x[i]);
}
void restated_atan2f_v2(const float* y, const float* x, float* out, long n) {
    for (long i = 0; i < n; i++) out[i] = urh_atan2f_v2(y[i], x[i]);
}
"""


def test_atan2f_bit_exact_vs_libm():
    hdr = os.path.join(ROOT, "urh_b200", "csrc", "fdlibm_atan2f.h")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "h.c")
        so = os.path.join(td, "h.so")
        open(src, "w").write(HARNESS % hdr)
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src, "-lm"])
        lib = ctypes.CDLL(so)
        libm = ctypes.CDLL("libm.so.6")
        libm.atan2f.restype = ctypes.c_float
        libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
        rng = np.random.default_rng(3)
        n = 400000
        parts_y, parts_x = [], []
        parts_y.append(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32))
        parts_x.append(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32))
        parts_y.append(rng.standard_normal(n).astype(np.float32))
        parts_x.append(rng.standard_normal(n).astype(np.float32))
        parts_y.append((rng.integers(-32768, 32768, n)).astype(np.float32))
        parts_x.append((rng.integers(-32768, 32768, n)).astype(np.float32))
        parts_y.append((rng.standard_normal(n) * 2.0 ** rng.integers(-40, 40, n)).astype(np.float32))
        parts_x.append((rng.standard_normal(n) * 2.0 ** rng.integers(-40, 40, n)).astype(np.float32))
        parts_y.append((rng.standard_normal(n) * 2.0 ** rng.integers(-120, 120, n)).astype(np.float32))
        parts_x.append((rng.standard_normal(n) * 2.0 ** rng.integers(-120, 120, n)).astype(np.float32))
        parts_y.append(rng.integers(-4, 5, n).astype(np.float32) * rng.choice([1.0, -1.0, 0.0, -0.0], n).astype(np.float32))
        parts_x.append(rng.integers(-4, 5, n).astype(np.float32) * rng.choice([1.0, -1.0, 1.0, -0.0], n).astype(np.float32))
        special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, 0.4375, 0.6875, 1.1875, 2.4375, 2.0**25, 2.0**-29], dtype=np.float32)
        sy, sx = np.meshgrid(special, special)
        parts_y.append(sy.ravel())
        parts_x.append(sx.ravel())
        y = np.ascontiguousarray(np.concatenate(parts_y))
        x = np.ascontiguousarray(np.concatenate(parts_x))
        out = np.empty_like(y)
        lib.restated_atan2f(y.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(y)))
        # libm reference, vectorised through a tiny C loop as well
        src2 = os.path.join(td, "m.c")
        so2 = os.path.join(td, "m.so")
        open(src2, "w").write("#include <math.h>\nvoid ref(const float*y,const float*x,float*o,long n){for(long i=0;i<n;i++)o[i]=atan2f(y[i],x[i]);}\n")
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-shared", "-fPIC", "-o", so2, src2, "-lm"])
        ref = np.empty_like(y)
        ctypes.CDLL(so2).ref(y.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), ref.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(y)))
        nan = np.isnan(ref)
        assert np.array_equal(np.isnan(out), nan)
        assert np.array_equal(out.view(np.uint32)[~nan], ref.view(np.uint32)[~nan])
        # the branch-light variant the FSK kernel actually calls
        out2 = np.empty_like(y)
        lib.restated_atan2f_v2(y.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), out2.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(y)))
        assert np.array_equal(np.isnan(out2), nan)
        assert np.array_equal(out2.view(np.uint32)[~nan], ref.view(np.uint32)[~nan])
