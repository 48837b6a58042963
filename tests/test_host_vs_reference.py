"""CPU: the host-side logic of the drop-in layer against the REFERENCE's own functions on randomized inputs (the reference
is imported through oracle/ref_loader.py; skipped where /root/reference does not exist).  These functions never touch the
GPU: pulses -> bits, plateau / bit-length bookkeeping of estimate(), modulator parameter preparation, filter design,
bit utilities."""
import array
import os

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src/urh"), reason="reference tree not present")


def _oracle_ppseq_to_bits(*a, **k):
    """the sequential CPU restatement of ProtocolAnalyzer._ppseq_to_bits lives in the test oracle, not in the product"""
    from oracle import oracle
    return oracle.ppseq_to_bits(*a, **k)


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    ns = ref_loader.load_python_layer()
    sf, ut, ai = ref_loader.load_kernels()
    ns.sf, ns.ut, ns.ai = sf, ut, ai
    return ns


def test_ppseq_to_bits_port(ref):
    from urh_b200.signalprocessing.ProtocolAnalyzer import ProtocolAnalyzer as PA
    rfun = ref.ProtocolAnalyzer(None)._ppseq_to_bits   # an instance method in the reference (ProtocolAnalyzer.py:323)
    rng = np.random.default_rng(5)
    for trial in range(300):
        bps = int(rng.choice([1, 2]))
        pt = int(rng.choice([8, 0, 2]))
        sps = int(rng.choice([1, 3, 10, 100]))
        k = int(rng.integers(1, 80))
        kinds = rng.integers(-1, 1 << bps, k)
        ns = np.where(rng.random(k) < 0.15, rng.integers(9, 30, k) * sps, rng.integers(0, 5 * sps + 1, k))
        rows = np.stack([kinds, ns], axis=1).astype(np.int64)
        wp = bool(trial % 2)
        mine = _oracle_ppseq_to_bits(rows, sps, bps, write_bit_sample_pos=wp, pause_threshold=pt)
        theirs = rfun(rows, sps, bps, write_bit_sample_pos=wp, pause_threshold=pt)
        assert [list(x) for x in mine[0]] == [list(x) for x in theirs[0]], trial
        assert list(mine[1]) == list(theirs[1]), trial
        assert [list(x) for x in mine[2]] == [list(x) for x in theirs[2]], trial


def test_plateau_bookkeeping(ref):
    from urh_b200.ainterpretation import AutoInterpretation as AI
    R = ref.AutoInterpretation
    rng = np.random.default_rng(9)
    for trial in range(300):
        n = int(rng.integers(2, 60))
        base = int(rng.choice([8, 40, 100, 300]))
        pl = (rng.integers(1, 6, n) * base + rng.integers(-base // 8 - 1, base // 8 + 2, n)).clip(1)
        if trial % 3 == 0:
            pl[rng.integers(0, n, max(1, n // 6))] = rng.integers(1, 4, max(1, n // 6))   # tiny glitches
        pl = pl.astype(np.uint64)
        assert AI.estimate_tolerance_from_plateau_lengths(pl) == R.estimate_tolerance_from_plateau_lengths(pl), trial
        for tol in (None, 0, 1, 3):
            assert list(AI.merge_plateau_lengths(pl, tolerance=tol)) == list(R.merge_plateau_lengths(pl, tolerance=tol)), (trial, tol)
        merged = R.merge_plateau_lengths(pl)
        if len(merged) >= 2:
            assert AI.get_bit_length_from_plateau_lengths(merged) == R.get_bit_length_from_plateau_lengths(merged), trial
        a, b = [int(v) for v in pl], [int(v) for v in pl]
        AI.round_plateau_lengths(a)       # in place
        R.round_plateau_lengths(b)
        assert a == b, trial
        assert AI.get_tolerant_greatest_common_divisor(list(pl)) == R.get_tolerant_greatest_common_divisor(list(pl)), trial
        vals = [int(v) for v in rng.integers(0, 6, n)]
        assert AI.get_most_frequent_value(vals) == R.get_most_frequent_value(vals)
        data = rng.standard_normal(n + 3) * 10 + 50
        assert AI.max_without_outliers(data) == R.max_without_outliers(data)
        assert AI.min_without_outliers(data) == R.min_without_outliers(data)


def test_cython_host_helpers(ref):
    from urh_b200.cythonext import auto_interpretation as cai, signal_functions as sf, util
    rng = np.random.default_rng(2)
    for trial in range(200):
        n = int(rng.integers(1, 80))
        pl = rng.integers(1, 400, n).astype(np.uint64)
        tol, mc = int(rng.integers(0, 12)), int(rng.integers(1, 40))
        assert list(cai.merge_plateaus(pl, tol, mc)) == list(np.asarray(ref.ai.merge_plateaus(pl, tol, mc))), trial
        assert list(cai.get_threshold_divisor_histogram(pl)) == list(np.asarray(ref.ai.get_threshold_divisor_histogram(pl))), trial
        if trial % 10 == 0:   # long tables with repeated values and zeros (a message of thousands of rounded plateaus)
            big = (rng.integers(0, 7, 3000) * int(rng.choice([10, 100, 300])) + (rng.integers(0, 3, 3000) if trial % 20 else 0)).astype(np.uint64)
            if big.max() == 0:
                big[0] = 5
            assert np.array_equal(cai.get_threshold_divisor_histogram(big), np.asarray(ref.ai.get_threshold_divisor_histogram(big))), trial
        bits = rng.integers(0, 2, int(rng.integers(0, 40))).astype(np.uint8)
        assert list(sf.get_oqpsk_bits(bits)) == list(np.asarray(ref.sf.get_oqpsk_bits(bits))), trial
        if len(bits):
            a, b = sorted(rng.integers(0, len(bits) + 1, 2))
            assert util.bit_array_to_number(bits, int(b), int(a)) == ref.ut.bit_array_to_number(bits, int(b), int(a))
    for sr, sps, bt, fw in ((2e6, 100, 0.5, 1.0), (1e6, 8, 0.3, 1.5), (250e3, 33, 1.0, 0.7)):
        mine = sf.gauss_fir(sr, sps, bt, fw)
        theirs = np.asarray(ref.sf.get_gauss_fir(sr, sps, bt, fw)) if hasattr(ref.sf, "get_gauss_fir") else None
        if theirs is not None:
            assert np.array_equal(mine, theirs)


def test_modulator_and_filter_host_logic(ref):
    from urh_b200.signalprocessing.Filter import Filter
    from urh_b200.signalprocessing.Modulator import Modulator
    for bw in (0.001, 0.04, 0.08, 0.42):
        assert Filter.get_filter_length_from_bandwidth(bw) == ref.Filter.get_filter_length_from_bandwidth(bw)
        N = Filter.get_filter_length_from_bandwidth(bw)
        assert Filter.get_bandwidth_from_filter_length(N) == ref.Filter.get_bandwidth_from_filter_length(N)
        if N < 2000:
            assert np.array_equal(Filter.design_windowed_sinc_lpf(0.1, bw), ref.Filter.design_windowed_sinc_lpf(0.1, bw))
            assert np.array_equal(Filter.design_windowed_sinc_bandpass(-0.1, 0.2, bw), ref.Filter.design_windowed_sinc_bandpass(-0.1, 0.2, bw))
    for mod in ("ASK", "FSK", "PSK", "GFSK", "OQPSK"):
        for bps in ((1, 2, 3) if mod != "OQPSK" else (2,)):
            m, r = Modulator("m"), ref.Modulator("m")
            for o in (m, r):
                o.modulation_type = mod
                o.bits_per_symbol = bps
                o.sample_rate = 2e6
            assert list(m.get_default_parameters()) == list(r.get_default_parameters()), (mod, bps)
            assert m.modulation_order == r.modulation_order and m.is_binary_modulation == r.is_binary_modulation
            assert (m.is_amplitude_based, m.is_frequency_based, m.is_phase_based) == (r.is_amplitude_based, r.is_frequency_based, r.is_phase_based)


def test_iq_array_host_logic(ref):
    from urh_b200.signalprocessing.IQArray import IQArray
    rng = np.random.default_rng(4)
    for dt in (np.int8, np.uint8, np.int16, np.uint16, np.float32):
        assert IQArray.min_max_for_dtype(dt) == ref.IQArray.min_max_for_dtype(dt)
    c = (rng.standard_normal(10) + 1j * rng.standard_normal(10)).astype(np.complex64)
    for arr in (c, c.astype(np.complex128), rng.standard_normal(20).astype(np.float32), rng.integers(-100, 100, (10, 2)).astype(np.int16),
                rng.integers(0, 255, 20).astype(np.uint8)):
        assert np.array_equal(IQArray.convert_array_to_iq(arr), ref.IQArray.convert_array_to_iq(arr))
        a, b = IQArray(arr), ref.IQArray(arr)
        assert a.num_samples == b.num_samples and a.dtype == b.dtype and a.minimum == b.minimum and a.maximum == b.maximum
        assert np.array_equal(a.real, b.real) and np.array_equal(a.imag, b.imag)
    for name in ("x.complex", "x.cs8", "x.complex16u", "x.cu16", "x.complex32s", "x.wav"):
        exp = {"x.complex": np.float32, "x.cs8": np.int8, "x.complex16u": np.uint8, "x.cu16": np.uint16, "x.complex32s": np.int16, "x.wav": np.float32}[name]
        assert IQArray._dtype_for_filename(name) == exp


def test_ring_buffer(ref):
    """util/RingBuffer.py:7-140: push / pop / wrap-around / clear on randomized traffic"""
    import importlib
    from urh_b200.util.RingBuffer import RingBuffer
    RRing = importlib.import_module("urh.util.RingBuffer").RingBuffer
    from urh_b200.signalprocessing.IQArray import IQArray
    rng = np.random.default_rng(6)
    for dtype in (np.float32, np.int8):
        mine, theirs = RingBuffer(size=64, dtype=dtype), RRing(size=64, dtype=dtype)
        for step in range(300):
            if rng.random() < 0.55:
                k = int(rng.integers(1, 40))
                vals = (rng.standard_normal((k, 2)) * 50).astype(dtype)
                assert mine.will_fit(k) == theirs.will_fit(k)
                if mine.will_fit(k):
                    mine.push(IQArray(vals.copy()))
                    theirs.push(ref.IQArray(vals.copy()))
            else:
                k = int(rng.integers(1, 50))
                even = bool(step % 2)
                a, b = mine.pop(k, ensure_even_length=even), theirs.pop(k, ensure_even_length=even)
                assert np.array_equal(np.asarray(a), np.asarray(b)), (dtype, step)
            assert (mine.left_index, mine.right_index, mine.space_left, mine.is_empty, len(mine)) == \
                   (theirs.left_index, theirs.right_index, theirs.space_left, theirs.is_empty, len(theirs))
            assert np.array_equal(np.asarray(mine.view_data), np.asarray(theirs.view_data))
            if step % 97 == 0:
                mine.clear()
                theirs.clear()


def test_modulator_prepares_the_same_kernel_call(ref, monkeypatch):
    """Modulator.modulate (Modulator.py:215-255): the arguments handed to modulate_c are the reference's (both kernels are
    replaced by recorders here, so no GPU and no Cython code runs)"""
    import importlib
    import urh_b200.signalprocessing.Modulator as mine_mod
    ref_mod = importlib.import_module("urh.signalprocessing.Modulator")
    calls = {"mine": [], "ref": []}

    def recorder(key):
        def fake(bits, sps, mod_type, parameters, bps, a, f, phi, sr, pause, start, dtype=np.float32, gauss_bt=0.5, filter_width=1.0):
            calls[key].append((list(bits), sps, mod_type, [float(p) for p in parameters], bps, float(a), float(f), float(phi), float(sr),
                               pause, start, np.dtype(dtype), float(gauss_bt), float(filter_width)))
            total = (len(bits) // bps) * sps + pause
            return np.zeros((total, 2), dtype=dtype)
        return fake

    monkeypatch.setattr(mine_mod.signal_functions, "modulate_c", recorder("mine"))
    monkeypatch.setattr(ref_mod.signal_functions, "modulate_c", recorder("ref"))
    rng = np.random.default_rng(12)
    for mod in ("ASK", "FSK", "PSK", "GFSK"):
        for trial in range(6):
            m, r = mine_mod.Modulator("t"), ref_mod.Modulator("t")
            bps = int(rng.choice([1, 2]))
            cfg = dict(modulation_type=mod, bits_per_symbol=bps, samples_per_symbol=int(rng.choice([8, 100])), sample_rate=float(rng.choice([1e6, 2e6])),
                       carrier_freq_hz=float(rng.choice([0.0, 20e3])), carrier_amplitude=float(rng.choice([1.0, 0.5])),
                       carrier_phase_deg=float(rng.choice([0.0, 45.0])), gauss_bt=0.5, gauss_filter_width=1.0)
            for o in (m, r):
                for k_, v in cfg.items():
                    setattr(o, k_, v)
                o.parameters = o.get_default_parameters()
            nbits = int(rng.integers(0, 12)) * bps
            data = [int(b) for b in rng.integers(0, 2, nbits)]
            payload = "".join(map(str, data)) if trial % 2 else list(data)
            pause, start = int(rng.integers(0, 50)), int(rng.integers(0, 1000))
            dtype = [None, np.int8, np.int16, np.float32][trial % 4]
            a = m.modulate(payload, pause=pause, start=start, dtype=dtype)
            b = r.modulate(payload, pause=pause, start=start, dtype=dtype)
            assert a.data.shape == b.data.shape and a.dtype == b.dtype
    assert len(calls["mine"]) == len(calls["ref"]) > 0
    assert calls["mine"] == calls["ref"]


def test_spectrogram_geometry(ref):
    """Spectrogram.py:84-103: hop size, bin counts and the number of STFT frames (the reference's frame count is the
    shape of its strided view; ours is computed up front to size the device buffers)"""
    from urh_b200.signalprocessing.Spectrogram import Spectrogram
    rng = np.random.default_rng(1)
    for trial in range(40):
        n = int(rng.integers(1, 5000))
        w = int(rng.choice([16, 64, 256, 1024]))
        ov = float(rng.choice([0.5, 0.0, 0.75, 0.3]))
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        a, b = Spectrogram(x, window_size=w, overlap_factor=ov), ref.Spectrogram(x, window_size=w, overlap_factor=ov)
        assert (a.hop_size, a.time_bins, a.freq_bins) == (b.hop_size, b.time_bins, b.freq_bins)
        assert a._num_frames(n) == b.stft(x).shape[0], (n, w, ov)


def test_merge_message_segments_for_ook(ref):
    from urh_b200.ainterpretation import AutoInterpretation as AI
    rng = np.random.default_rng(21)
    assert AI.merge_message_segments_for_ook([]) == ref.AutoInterpretation.merge_message_segments_for_ook([])
    for trial in range(300):
        k = int(rng.integers(1, 25))
        pos = 0
        segs = []
        pulse = int(rng.choice([20, 100, 400]))
        for _ in range(k):
            pos += int(rng.choice([pulse // 2, pulse, 3 * pulse, 9 * pulse, 40 * pulse])) + int(rng.integers(0, 5))
            length = int(rng.integers(1, 4)) * pulse + int(rng.integers(0, 7))
            segs.append((pos, pos + length))
            pos += length
        assert AI.merge_message_segments_for_ook(list(segs)) == ref.AutoInterpretation.merge_message_segments_for_ook(list(segs)), trial


def test_noise_level_decision_from_chunk_statistics(ref):
    """detect_noise_level (AutoInterpretation.py:60-91): the device only delivers (sum, max) of the 100 end-aligned
    chunks; here they come from numpy, the decision logic is ours, the expected value is the reference's."""
    from urh_b200.ainterpretation import AutoInterpretation as AI
    rng = np.random.default_rng(14)
    for trial in range(200):
        n = int(rng.integers(4, 40000))
        dtype = np.float64 if trial % 2 else np.float32
        mags = np.abs(rng.standard_normal(n) * 0.01)
        kind = trial % 5
        if kind < 3:
            a = int(rng.integers(0, n))
            mags[a: a + n // 3] += rng.uniform(0.3, 1.0)          # a burst
        elif kind == 3:
            mags += 0.5                                            # signal everywhere: chunk means nearly equal -> 0
        else:
            mags[:] = 0.0
        mags = mags.astype(dtype)
        chunksize, nchunks = AI._chunking(n)
        tail = mags[n - nchunks * chunksize:].reshape(nchunks, chunksize)   # chunks are taken from the end backwards
        sums = tail.astype(np.float64).sum(axis=1)
        maxs = tail.max(axis=1).astype(np.float64)
        got = AI._noise_from_chunk_stats(n, chunksize, sums, maxs, dtype)
        assert got == ref.AutoInterpretation.detect_noise_level(mags), (trial, n)


def test_oracle_convert_iq_is_the_references_convert_to(ref, oracle):
    """closes the chain for the format conversions: reference IQArray.convert_to == oracle.convert_iq (here) == convert.cu
    (tests/test_gpu_objects.py::test_convert_to_all_pairs)"""
    rng = np.random.default_rng(5)
    types = [np.int8, np.uint8, np.int16, np.uint16, np.float32]
    for src in types:
        if src == np.float32:
            x = np.concatenate([rng.uniform(-1, 1, 4000), [-1.0, 1.0, 0.0, -0.0, 0.999999, -0.999999]]).astype(np.float32)
        else:
            info = np.iinfo(src)
            x = np.concatenate([rng.integers(info.min, info.max + 1, 4000), [info.min, info.max, 0, 1]]).astype(src)
        x = np.ascontiguousarray(x.reshape(-1, 2))
        for dst in types:
            a = oracle.convert_iq(x, dst)
            b = ref.IQArray(x).convert_to(dst)
            assert a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8)), (src, dst)
