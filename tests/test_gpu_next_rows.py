"""GPU: the rows SURVEY 8f marks "next": the live-sniffer demodulation loop as a streaming client (f-3), per-bit frequency
estimation and the spectrogram's BGRA colormap look-up (f-4)."""
import array

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _modulated_messages(rng, nmsg, sps, pause):
    from urh_b200.signalprocessing.Modulator import Modulator

    m = Modulator("sniff")
    m.modulation_type = "FSK"
    m.parameters = array.array("f", [-20e3, 20e3])
    m.samples_per_symbol = sps
    m.sample_rate = 1e6
    m.carrier_freq_hz = 0
    msgs, parts = [], []
    for _ in range(nmsg):
        bits = [1, 0, 1, 0, 1, 0, 1, 0] + [int(b) for b in rng.integers(0, 2, int(rng.integers(24, 80)))] + [1]
        msgs.append("".join(map(str, bits)))
        parts.append(m.modulate(bits, pause=pause).data)
    iq = np.concatenate(parts)
    iq = iq + 0.005 * rng.standard_normal(iq.shape).astype(np.float32)
    return msgs, iq.astype(np.float32).view(np.complex64).reshape(-1)


@pytest.mark.parametrize("chunk", [4096, 10_000, 123_457])
def test_sniffer_stream_recovers_the_messages(chunk):
    """ProtocolSniffer.feed == the reference's __demodulate_data (ProtocolSniffer.py:204-283): chunks in, messages out.  Shaped like
    the reference's tests/test_protocol_sniffer.py:33-95 (modulated messages with pauses arrive in pieces; the sniffed bit strings
    are the sent ones)."""
    from urh_b200.signalprocessing.ProtocolSniffer import ProtocolSniffer

    rng = np.random.default_rng(chunk)
    sps = 100
    sent, samples = _modulated_messages(rng, 6, sps, pause=20 * sps)
    sniffer = ProtocolSniffer(sps, 0.0, 0.1, 0.05, 5, "FSK", 1, sample_rate=1e6)
    for a in range(0, len(samples), chunk):
        sniffer.feed(samples[a:a + chunk])
    sniffer.flush()
    got = [m.plain_bits_str for m in sniffer.messages]
    # a message's trailing zeros belong to the pause (FSK: 0 = the lower tone is only known from a pulse): compare up to them
    assert len(got) == len(sent), (got, sent)
    for g, s in zip(got, sent):
        assert g.rstrip("0") == s.rstrip("0")
    stamps = [m.timestamp for m in sniffer.messages]
    assert all(b > a for a, b in zip(stamps, stamps[1:]))


def test_sniffer_adaptive_noise_and_automatic_center():
    from urh_b200.signalprocessing.ProtocolSniffer import ProtocolSniffer

    rng = np.random.default_rng(3)
    sps = 100
    sent, samples = _modulated_messages(rng, 3, sps, pause=30 * sps)
    sniffer = ProtocolSniffer(sps, 0.3, 0.1, 0.5, 5, "FSK", 1)   # wrong center and a far too high noise level to start with
    sniffer.adaptive_noise = True
    sniffer.automatic_center = True
    silence = (0.005 * (rng.standard_normal(50_000) + 1j * rng.standard_normal(50_000))).astype(np.complex64)
    for _ in range(40):   # the threshold decays towards the noise peaks (ProtocolSniffer.py:216-220)
        sniffer.feed(silence[:5000])
    assert sniffer.signal.noise_threshold < 0.06
    for a in range(0, len(samples), 8192):
        sniffer.feed(samples[a:a + 8192])
    sniffer.flush()
    assert [m.plain_bits_str.rstrip("0") for m in sniffer.messages] == [s.rstrip("0") for s in sent]
    assert abs(sniffer.signal.center) < 0.05


def test_estimate_frequency_for_bits():
    """ProtocolAnalyzer.estimate_frequency_for_one / _zero (ProtocolAnalyzer.py:570-632) == the reference's loop with the numpy FFT"""
    from urh_b200.signalprocessing.IQArray import IQArray
    from urh_b200.signalprocessing.ProtocolAnalyzer import ProtocolAnalyzer
    from urh_b200.signalprocessing.Signal import Signal

    g = load_golden("capture_fsk")
    s = Signal("", "t")
    s.iq_array = IQArray(g["iq"])
    s.noise_threshold = float(g["noise"])
    s.modulation_type = "FSK"
    s.samples_per_symbol, s.center, s.tolerance = 100, 0.0, 5
    pa = ProtocolAnalyzer(s)
    pa.get_protocol_from_signal()
    data = s.iq_array.as_complex64()

    def ref(bit, nbits=42):
        freqs = []
        for i, msg in enumerate(pa.messages):
            for j, b in enumerate(msg.plain_bits):
                if b == bit:
                    start, num = pa.get_samplepos_of_bitseq(i, j, i, j + 1, False)
                    length = 2 ** int(np.log2(num))
                    w = np.fft.fft(data[start:start + length])
                    freqs.append(abs(np.fft.fftfreq(len(w))[np.argmax(np.abs(w))] * 1e6))
                    if len(freqs) == nbits:
                        return np.mean(freqs)
        return np.mean(freqs) if freqs else 0

    one, zero = pa.estimate_frequency_for_one(1e6), pa.estimate_frequency_for_zero(1e6)
    assert one == ref(True) and zero == ref(False)
    assert one != zero and pa.estimate_frequency_for_one(1e6, nbits=0) == 0


def test_bgra_lookup_matches_numpy():
    from urh_b200.signalprocessing.Spectrogram import Spectrogram

    rng = np.random.default_rng(9)
    colormap = rng.integers(0, 256, (256, 4), dtype=np.uint8)
    data = (rng.standard_normal((333, 257)) * 40 - 60).astype(np.float32)
    data[0, 0], data[1, 1], data[2, 2] = np.nan, np.inf, -np.inf

    def ref(d, cm, mn, mx, normalize=True):   # Spectrogram.py:192-206
        with np.errstate(all="ignore"):
            nv = (len(cm) - 1) * ((d.T - mn) / (mx - mn)) if normalize else d.T
            return np.take(cm, nv.astype(int), axis=0, mode="clip")

    for mn, mx in ((-140, 10), (-100.5, -20.25)):
        got = Spectrogram.apply_bgra_lookup(data, colormap, mn, mx)
        want = ref(data, colormap, mn, mx)
        assert got.shape == want.shape == (257, 333, 4)
        assert np.array_equal(got, want)
    idx = rng.integers(-5, 300, (64, 50)).astype(np.float32)
    assert np.array_equal(Spectrogram.apply_bgra_lookup(idx, colormap, normalize=False), ref(idx, colormap, 0, 1, normalize=False))
    with pytest.raises(ValueError):
        Spectrogram.apply_bgra_lookup(data, colormap)
