"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only where /root/reference exists (the build container).  It imports the reference's own Python
layer (Signal, ProtocolAnalyzer, Modulator, Filter, Spectrogram, AutoInterpretation) on top of the
reference's compiled Cython kernels (oracle/_ref, see oracle/build_ref.py + oracle/ref_loader.py) and
stores inputs + outputs as compressed .npz files.  The inputs are (slices of) the reference's own test
captures in tests/data plus seeded synthetic signals.

    python tests/golden/make_golden.py
"""
import array
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in kw.items()})


def main():
    ns = ref_loader.load_python_layer()
    sf, ut, ai = ref_loader.load_kernels()
    AI = ns.AutoInterpretation

    # ---- 1. captures: demod / digitize / noise / segmentation / center ------------------------------------
    captures = [
        # file, modulation, sps, center, tolerance, noise (None = auto), bits_per_symbol, spacing, max_samples
        ("fsk.complex", "FSK", 100, 0.0, 5, None, 1, 1.0, None),
        ("ask.complex", "ASK", 295, 0.0219, 5, None, 1, 1.0, None),
        ("ask_short.complex", "ASK", 16, 0.13, 0, 0.0299, 1, 1.0, None),
        ("psk_gen_noisy.complex", "PSK", 300, 0.0, 10, 0.0, 1, 1.0, None),
        ("enocean.complex", "ASK", 40, 0.04, 1, None, 1, 1.0, None),
        ("FSK10.complex", "FSK", 10, 0.0, 1, None, 1, 1.0, None),
        ("homematic.complex32s", "FSK", 100, 0.0, 5, None, 1, 1.0, 60000),
        ("esaver.complex16s", "ASK", 100, 0.02, 5, None, 1, 1.0, 120000),
        ("two_participants.complex16s", "FSK", 100, 0.0, 5, None, 1, 1.0, 120000),
    ]
    for fname, mod, sps, center, tol, noise, bps, spacing, max_samples in captures:
        path = os.path.join(ns.data_dir, fname)
        sig = ns.Signal(path, "golden")
        if max_samples is not None and sig.num_samples > max_samples:
            iq = np.ascontiguousarray(sig.iq_array.data[:max_samples])
            sig2 = ns.Signal("", "golden")
            sig2.iq_array = ns.IQArray(iq)
            sig2.noise_threshold = AI.detect_noise_level(sig2.iq_array.magnitudes)
            sig = sig2
        sig.modulation_type = mod
        sig.samples_per_symbol = sps
        sig.center = center
        sig.tolerance = tol
        sig.bits_per_symbol = bps
        sig.center_spacing = spacing
        auto_noise = sig.noise_threshold
        if noise is not None:
            sig.noise_threshold = noise
        iq = np.ascontiguousarray(sig.iq_array.data)
        out = {"iq": iq, "auto_noise": np.float64(auto_noise), "noise": np.float64(sig.noise_threshold)}
        for m in ("ASK", "FSK", "PSK"):
            q = sf.afp_demod(iq, sig.noise_threshold, m, 2, 0.1)
            q = np.array(q)
            if m == "PSK":
                q[0] = 0.0  # uninitialised in the reference (np.empty)
            out["qad_" + m] = q
        q4 = np.array(sf.afp_demod(iq, sig.noise_threshold, "PSK", 4, 0.1))
        q4[0] = 0.0
        out["qad_PSK4"] = q4
        qad = out["qad_" + mod]
        for t in (0, 1, tol, 17):
            out["pulses_tol%d" % t] = np.array(sf.grab_pulse_lens(qad, center, t, mod, sps, bps, spacing))
        out["pulses_bps2"] = np.array(sf.grab_pulse_lens(qad, center, tol, mod, sps, 2, 0.1))
        pa = ns.ProtocolAnalyzer(sig)
        pa.get_protocol_from_signal()
        bits = pa.plain_bits_str
        mags = ut.get_magnitudes(iq)
        out["mag_sum"] = np.float64(mags.sum())
        out["mag_head"] = mags[:64]
        out["segments"] = np.array(AI.segment_messages_from_magnitudes(mags, sig.noise_threshold), dtype=np.int64).reshape(-1, 2)
        c = AI.detect_center(qad)
        out["detect_center"] = np.float64(np.nan if c is None else c)
        est = AI.estimate(sig.iq_array)
        meta = dict(file=fname, mod=mod, sps=sps, center=center, tol=tol, bps=bps, spacing=spacing, bits=bits,
                    estimate=None if est is None else {k: (float(v) if not isinstance(v, str) else v) for k, v in est.items()})
        out["meta"] = np.array(json.dumps(meta))
        save("capture_" + fname.split(".")[0], **out)

    # ---- 2. modulator -------------------------------------------------------------------------------------
    rng = np.random.default_rng(1234)
    mods = {}
    bits = rng.integers(0, 2, 96).astype(np.uint8)
    cases = [
        ("ask", "ASK", [0, 100], 1, np.float32), ("ask_i8", "ASK", [0, 100], 1, np.int8),
        ("fsk", "FSK", [-10e3, 10e3], 1, np.float32), ("fsk4", "FSK", [-20e3, -10e3, 10e3, 20e3], 2, np.float32),
        ("fsk_i16", "FSK", [-10e3, 10e3], 1, np.int16),
        ("psk", "PSK", [-90, 90], 1, np.float32), ("psk4", "PSK", [-135, -45, 45, 135], 2, np.float32),
        ("oqpsk", "OQPSK", [-135, -45, 45, 135], 2, np.float32),
        ("gfsk", "GFSK", [-10e3, 10e3], 1, np.float32), ("gfsk_i8", "GFSK", [-10e3, 10e3], 1, np.int8),
    ]
    for name, mt, params, bps, dt in cases:
        m = ns.Modulator("golden")
        m.modulation_type = mt
        m.bits_per_symbol = bps
        m.parameters = array.array("f", params)
        m.samples_per_symbol = 50
        m.sample_rate = 1e6
        m.carrier_freq_hz = 40e3
        m.carrier_phase_deg = 30
        res = m.modulate(list(map(int, bits)), pause=77, start=0, dtype=dt).data
        mods["mod_" + name] = np.array(res)
        mods["mod_" + name + "_start5"] = np.array(m.modulate(list(map(int, bits[:32])), pause=3, start=5, dtype=dt).data)
    mods["bits"] = bits
    save("modulator", **mods)

    # ---- 3. filters / spectrogram -------------------------------------------------------------------------
    x = (rng.standard_normal(5000) + 1j * rng.standard_normal(5000)).astype(np.complex64)
    taps = (rng.standard_normal(17) + 1j * rng.standard_normal(17)).astype(np.complex64)
    f = {"x": x, "taps": taps, "fir": sf.fir_filter(x, taps)}
    ma = np.array([0.1] * 10, dtype=np.complex64)
    f["fir_ma10"] = sf.fir_filter(x, ma)
    f["kat_in"] = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 42], dtype=np.complex64)
    f["kat_out"] = ns.Filter([0.25, 0.25, 0.25, 0.25]).apply_fir_filter(f["kat_in"].flatten())
    f["bandpass_taps"] = ns.Filter.design_windowed_sinc_bandpass(0.03, 0.07, 0.04)
    f["bandpass_direct"] = ns.Filter.apply_bandpass_filter(x[:300], 0.03, 0.07, 0.2)   # short taps -> direct path
    f["bandpass_fft"] = ns.Filter.apply_bandpass_filter(x, 0.03, 0.07, 0.04)           # 101 taps -> FFT path
    iqx = x.view(np.float32).reshape(-1, 2)
    f["dc"] = ns.Filter([], ns.FilterType.dc_correction).work(iqx)
    sp = ns.Spectrogram(x)
    f["stft"] = sp.stft(x).astype(np.complex64)
    spec = np.fft.fftshift(sp.stft(x), axes=(1,))
    f["spec_db"] = np.fliplr(ut.arr2decibel(spec.astype(np.complex64)))
    f["short_db"] = np.fliplr(ut.arr2decibel(np.fft.fftshift(sp.stft(x[:300]), axes=(1,)).astype(np.complex64)))
    save("filters", **f)


if __name__ == "__main__":
    main()
