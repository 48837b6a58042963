"""GPU parity: pulse table -> bits (bits.cu) vs the host restatement of ProtocolAnalyzer._ppseq_to_bits, which the
reference's own demodulation tests pin (tests/test_gpu_objects.py compares it with golden bit strings)."""
import numpy as np
import pytest

from conftest import CAPTURES, load_golden

pytestmark = pytest.mark.gpu


def _oracle_ppseq_to_bits(*a, **k):
    """the sequential CPU restatement of ProtocolAnalyzer._ppseq_to_bits lives in the test oracle, not in the product"""
    from oracle import oracle
    return oracle.ppseq_to_bits(*a, **k)


def both(ppseq, sps, bps, pt, write_pos=True):
    from urh_b200.signalprocessing.ProtocolAnalyzer import ProtocolAnalyzer as PA
    host = _oracle_ppseq_to_bits(ppseq, sps, bps, write_bit_sample_pos=write_pos, pause_threshold=pt)
    dev = PA._ppseq_to_bits_device(ppseq, sps, bps, write_bit_sample_pos=write_pos, pause_threshold=pt)
    return host, dev


def assert_same(host, dev):
    hb, hp, hpos = host
    db, dp, dpos = dev
    assert len(hb) == len(db)
    for a, b in zip(hb, db):
        assert a.tobytes() == b.tobytes()
    assert list(hp) == list(dp)
    assert len(hpos) == len(dpos)
    for a, b in zip(hpos, dpos):
        assert list(a) == list(b)


def random_table(rng, k, sps, kinds, start_pause):
    rows = np.empty((k, 2), np.int64)
    last = None
    for i in range(k):
        while True:
            kind = int(rng.choice(kinds))
            if kind != last:
                break
        last = kind
        r = rng.random()
        if r < 0.6:
            ns = int(rng.integers(1, 6)) * sps + int(rng.integers(-sps // 3, sps // 3 + 1))
        elif r < 0.8:
            ns = int(rng.integers(0, sps))              # fractions around the 0.5 rounding rule
        elif r < 0.9:
            ns = sps // 2 + int(rng.integers(-1, 2))   # exactly at the rule
        else:
            ns = int(rng.integers(9, 40)) * sps         # long: a message separator when it is a pause
        rows[i] = (kind, max(ns, 0))
    if start_pause:
        rows[0, 0] = -1
    return rows


@pytest.mark.parametrize("bps", [1, 2, 3])
@pytest.mark.parametrize("pt", [8, 0, 1])
def test_random_tables(bps, pt):
    rng = np.random.default_rng(100 * bps + pt)
    kinds = [-1] + list(range(1 << bps))
    for trial in range(25):
        sps = int(rng.choice([1, 2, 7, 10, 100, 333]))
        k = int(rng.integers(1, 300))
        rows = random_table(rng, k, sps, kinds, start_pause=bool(trial % 2))
        assert_same(*both(rows, sps, bps, pt, write_pos=bool(trial % 3)))


def test_edge_tables():
    E = lambda *r: np.array(r, np.int64).reshape(-1, 2)
    cases = [
        E((-1, 1000)),                                   # only a pause
        E((1, 100)),                                     # one data row, no pause
        E((1, 100), (-1, 50)),                           # ends with a short pause
        E((1, 100), (-1, 5000)),                         # ends with a long pause
        E((-1, 5000), (1, 100), (0, 200), (-1, 5000), (-1, 5000), (1, 30)),   # double separator, last symbol rounds to 0
        E((-1, 300), (-1, 5000), (1, 100)),              # zeros dropped by a separator before any data
        E((0, 40), (1, 40), (0, 40)),                    # every row rounds to 0 symbols: no message
        E((1, 151), (0, 150), (1, 149)),                 # the 0.5 rule
    ]
    for rows in cases:
        for pt in (8, 0):
            assert_same(*both(rows, 100, 1, pt))


@pytest.mark.parametrize("name", CAPTURES)
def test_golden_captures(name):
    g = load_golden("capture_" + name)
    m = g["meta"]
    keys = [k for k in g.keys() if k.startswith("pulses_tol")]
    assert keys
    for key in keys:
        assert_same(*both(g[key], int(m["sps"]), 1, 8))


def test_table_left_on_device_and_scale():
    """no host round trip for the rows: digitize, then bits from the table in the context; 2^22 samples ~ 2e4 rows"""
    from conftest import synth_fsk
    from urh_b200.cythonext import signal_functions as sf
    from urh_b200.signalprocessing.ProtocolAnalyzer import ProtocolAnalyzer as PA
    n = 1 << 22
    iq = synth_fsk(n, sps=100, seed=3, gap_every=400_000)
    qad, rows = sf.demod_digitize(iq, 0.05, "FSK", 0.0, 5, 100)
    bits, off, pauses, pos = sf.ppseq_to_bits(len(rows), 100, 1)
    hb, hp, hpos = _oracle_ppseq_to_bits(rows, 100, 1)
    assert len(hb) == len(pauses) and list(hp) == list(pauses)
    for m in range(len(hb)):
        assert hb[m].tobytes() == bits[off[m]:off[m + 1]].tobytes()
        assert list(hpos[m]) == list(pos[off[m] + 2 * m: min(off[m + 1] + 2 * m + 2, len(pos))])
