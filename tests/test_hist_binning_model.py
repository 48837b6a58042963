"""CPU model of the histogram binning in center.cu (HistBins::bin_of<FAST>): float thresholds ru(edge) make the float
comparison exact, and the one-look-up bin guess rn((f - hmin) / hstep) is never more than one bin high under the condition
the host checks (|edge| / hstep < 2^20).  The model mirrors the device arithmetic (float32 FFMA, round-to-nearest-even,
one table look-up) in numpy and is compared with np.histogram's own bin assignment on double edges."""
import numpy as np


def ru(x):
    """smallest float32 >= x (x: float64 array)"""
    f = x.astype(np.float32)
    low = f.astype(np.float64) < x
    return np.where(low, np.nextafter(f, np.float32(np.inf)), f).astype(np.float32)


def rd(x):
    f = x.astype(np.float32)
    high = f.astype(np.float64) > x
    return np.where(high, np.nextafter(f, np.float32(-np.inf)), f).astype(np.float32)


def device_bins(f, hmin, hstep, nbins):
    """HistBins::bin_of<true> for float32 samples f; -1 = not counted"""
    k = np.arange(nbins + 1, dtype=np.float64)
    edges = hmin + k * hstep                       # np.arange's own formula: one rounded product, one rounded sum
    fe = ru(edges)
    f_hi = rd(edges[-1:])[0]
    f_min = max(fe[0], np.nextafter(np.float32(-4.0), np.float32(0.0)))
    scale = np.float32(1.0 / hstep)
    off = np.float32(-(np.float64(fe[0]) * np.float64(scale)))          # -fe[0] * scale in float32
    valid = (f >= f_min) & (f <= f_hi)
    t = (f.astype(np.float64) * np.float64(scale) + np.float64(off)).astype(np.float32)   # FFMA (double-rounding aside)
    with np.errstate(invalid="ignore"):
        r = np.rint(t.astype(np.float64)).astype(np.int64)
    r = np.clip(r, 0, nbins)
    kk = r - (f < fe[r]).astype(np.int64)
    kk = np.minimum(kk, nbins - 1)
    return np.where(valid, kk, -1), edges


def reference_bins(f, edges):
    """np.histogram's assignment on the double edges: right-open bins, the last one closed; -1 outside"""
    a = f.astype(np.float64)
    k = np.searchsorted(edges, a, side="right") - 1
    k = np.where(a == edges[-1], len(edges) - 2, k)
    inside = (a >= edges[0]) & (a <= edges[-1]) & (f > np.float32(-4.0))
    return np.where(inside, k, -1)


def test_fast_binning_is_exact_under_the_host_condition():
    rng = np.random.default_rng(17)
    worst = 0
    for trial in range(400):
        nbins = int(rng.choice([1, 2, 7, 64, 300, 1593, 6000]))
        span = float(10 ** rng.uniform(-3, 0.8))
        hmin = float(rng.uniform(-3.9, 3.0)) if trial % 5 else float(rng.uniform(-3.9, -3.0))
        hstep = span / nbins
        if max(abs(hmin), abs(hmin + nbins * hstep)) / hstep >= 2 ** 20:
            continue   # the host launches the loop-based variant there
        # samples: uniform over a slightly wider range, the edges themselves and their float neighbours, NOISE sentinels
        k = np.arange(nbins + 1, dtype=np.float64)
        edges = hmin + k * hstep
        e32 = edges.astype(np.float32)
        f = np.concatenate([
            rng.uniform(hmin - 2 * hstep, hmin + (nbins + 2) * hstep, 20000).astype(np.float32),
            e32, np.nextafter(e32, np.float32(np.inf)), np.nextafter(e32, np.float32(-np.inf)),
            np.full(10, -4.0, np.float32), np.array([np.nan, np.inf, -np.inf, 0.0, -0.0], np.float32)])
        got, edges = device_bins(f, hmin, hstep, nbins)
        ref = reference_bins(f, edges)
        bad = np.nonzero(got != ref)[0]
        assert len(bad) == 0, (trial, nbins, hmin, hstep, f[bad][:5], got[bad][:5], ref[bad][:5])
        worst = max(worst, nbins)
    assert worst >= 6000
