"""Host half of detect_center (no GPU): vectorised peak picking == the reference's per-bin loop
(AutoInterpretation.py:213-240), rank window and statistics bookkeeping."""
import numpy as np

from urh_b200.ainterpretation import AutoInterpretation as AI


def loop_pick(y, edges):
    """literal restatement of the reference loop"""
    num_values = len(y)
    most_common_levels = []
    window_size = max(2, int(0.05 * num_values) + 1)
    for index in np.argsort(y)[::-1]:
        is_local_maximum = True
        for i in range(1, window_size):
            right = y[index + i] if index + i < num_values else 0
            left = y[index - i] if index - i >= 0 else 0
            if not (y[index] > right and y[index] > left):
                is_local_maximum = False
                break
        if is_local_maximum:
            most_common_levels.append(edges[index])
        if len(most_common_levels) == 2:
            break
    if len(most_common_levels) == 0:
        return None
    return np.mean(most_common_levels)


def test_peak_picking_matches_loop():
    rng = np.random.default_rng(11)
    for trial in range(300):
        nbins = int(rng.integers(1, 400))
        kind = trial % 4
        if kind == 0:
            y = rng.integers(0, 50, nbins)
        elif kind == 1:
            y = np.zeros(nbins, dtype=np.int64)
            y[rng.integers(0, nbins, 3)] = rng.integers(1, 1000, 3)
        elif kind == 2:
            g = np.arange(nbins)
            y = (1000 * np.exp(-0.5 * ((g - nbins * 0.3) / 3.0) ** 2) + 800 * np.exp(-0.5 * ((g - nbins * 0.7) / 3.0) ** 2)).astype(np.int64)
            y += rng.integers(0, 3, nbins)
        else:
            y = np.full(nbins, 7, dtype=np.int64)  # plateau: no strict maximum
        edges = np.arange(nbins + 1) * 0.01 - 1.0
        a, b = AI.pick_center_from_histogram(y.astype(np.int64), edges), loop_pick(y.astype(np.int64), edges)
        assert (a is None) == (b is None)
        if a is not None:
            assert a == b


def test_rank_window_and_stats():
    assert AI.center_rank_window(1000) == (50, 950)
    assert AI.center_rank_window(1000, 100) == (50, 150)
    assert AI.center_rank_window(10, None) == (0, 9)
    assert AI.center_rank_window(0) == (0, 0)
    st = AI.center_stats_from_window(100, 5, 95, np.array([90.0, -1.0, 2.0, 45.0, 100.0]))
    assert st[3] == -1.0 and st[4] == 2.0 and st[5] == 0.5 and abs(st[6] - (100.0 - 90 * 0.25) / 90) < 1e-15
    empty = AI.center_stats_from_window(100, 5, 95, np.zeros(5))
    assert AI.center_bin_edges(empty) is None
    const = AI.center_stats_from_window(100, 5, 95, np.array([90.0, 1.0, 1.0, 90.0, 90.0]))
    assert AI.center_bin_edges(const) is None  # zero variance: arange raises -> no center


def test_fsk_peak_test_from_spectrum_features_equals_reference_rule():
    """detect_modulation's FSK rule (AutoInterpretation.py:196-206) evaluated from the three spectrum features the device
    returns (arg-max, best bin >= 10 away, the 19 bins around the arg-max) == the rule on the full spectrum."""
    rng = np.random.default_rng(3)
    for trial in range(2000):
        P = int(rng.choice([32, 64, 256, 1024]))
        kind = trial % 5
        fft = rng.uniform(0, 50, P)
        if kind == 1:
            fft[rng.integers(0, P)] += 500                      # one peak
        elif kind == 2:
            a = int(rng.integers(0, P))
            fft[a] += 500
            fft[(a + int(rng.integers(10, P - 10))) % P] += float(rng.choice([80, 150, 400]))   # a second peak far away
        elif kind == 3:
            a = int(rng.integers(0, P))
            fft[max(0, a - 9):a + 10] += rng.uniform(200, 900, len(fft[max(0, a - 9):a + 10]))  # a broad peak: 19 big bins
            fft[(a + P // 2) % P] += float(rng.choice([0, 120, 300, 950]))
        elif kind == 4:
            fft = rng.uniform(90, 110, P)                       # everything around the threshold
        fft = fft.astype(np.float32)
        ten = np.argsort(fft)[::-1][0:10]
        ref = bool(any(abs(i - ten[0]) >= 10 and fft[i] >= 100 for i in ten))
        g = int(np.argmax(fft))
        assert g == ten[0]
        spec = np.full(23, -1.0)
        spec[0], spec[1] = g, fft[g]
        far = [i for i in range(P) if abs(i - g) >= 10]
        if far:
            f = max(far, key=lambda i: fft[i])
            spec[2], spec[3] = f, fft[f]
        for j in range(19):
            i = g - 9 + j
            if 0 <= i < P:
                spec[4 + j] = fft[i]
        assert AI._fsk_peak_test(spec) == ref, (trial, P)
