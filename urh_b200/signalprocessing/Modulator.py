"""``Modulator`` forward path (reference: src/urh/signalprocessing/Modulator.py:17-276).  Parameter object +
``modulate`` -> CUDA (modulate.cu).  XML persistence and the Qt preview scenes of the reference class are GUI /
project plumbing and out of scope."""
import array
import math

import numpy as np

from .. import settings
from ..cythonext import signal_functions
from .IQArray import IQArray


class Modulator(object):
    FORCE_DTYPE = None

    MODULATION_TYPES = ["ASK", "FSK", "PSK", "GFSK", "OQPSK"]
    MODULATION_TYPES_VERBOSE = {
        "ASK": "Amplitude Shift Keying (ASK)",
        "FSK": "Frequency Shift Keying (FSK)",
        "PSK": "Phase Shift Keying (PSK)",
        "OQPSK": "Offset Quadrature Phase Shift Keying (OQPSK)",
        "GFSK": "Gaussian Frequeny Shift Keying (GFSK)",
    }

    def __init__(self, name: str):
        self.carrier_freq_hz = 40 * 10 ** 3
        self.carrier_amplitude = 1
        self.carrier_phase_deg = 0
        self.data = [True, False, True, False]
        self.samples_per_symbol = 100
        self.default_sample_rate = 10 ** 6
        self.__sample_rate = None
        self.__modulation_type = "ASK"
        self.__bits_per_symbol = 1
        self.name = name
        self.gauss_bt = 0.5
        self.gauss_filter_width = 1
        self.parameters = array.array("f", [0, 100])  # Freq, Amplitude (0..100 %) or Phase (0..360)

    def __eq__(self, other):
        keys = ("carrier_freq_hz", "carrier_amplitude", "carrier_phase_deg", "name", "modulation_type",
                "samples_per_symbol", "bits_per_symbol", "sample_rate", "parameters")
        return all(getattr(self, k) == getattr(other, k) for k in keys)

    @staticmethod
    def get_dtype():
        if Modulator.FORCE_DTYPE is not None:
            return Modulator.FORCE_DTYPE
        return {"int8": np.int8, "int16": np.int16}.get(settings.read("modulation_dtype", "float32", str), np.float32)

    @property
    def modulation_type(self) -> str:
        return self.__modulation_type

    @modulation_type.setter
    def modulation_type(self, value):
        try:
            self.__modulation_type = self.MODULATION_TYPES[int(value)]  # legacy integer index
        except (ValueError, IndexError):
            self.__modulation_type = value

    @property
    def is_binary_modulation(self):
        return self.bits_per_symbol == 1

    @property
    def is_amplitude_based(self):
        return "ASK" in self.modulation_type

    @property
    def is_frequency_based(self):
        return "FSK" in self.modulation_type

    @property
    def is_phase_based(self):
        return "PSK" in self.modulation_type

    @property
    def bits_per_symbol(self):
        return self.__bits_per_symbol

    @bits_per_symbol.setter
    def bits_per_symbol(self, value):
        value = int(value)
        if value != self.bits_per_symbol:
            self.__bits_per_symbol = value
            self.parameters = array.array("f", [0] * self.modulation_order)

    @property
    def modulation_order(self):
        return 2 ** self.bits_per_symbol

    @property
    def sample_rate(self):
        return self.__sample_rate if self.__sample_rate is not None else self.default_sample_rate

    @sample_rate.setter
    def sample_rate(self, value):
        self.__sample_rate = value

    @property
    def carrier_phase_rad(self):
        return self.carrier_phase_deg * (np.pi / 180)

    def _prepared(self, data, dtype):
        if isinstance(data, str):
            data = array.array("B", map(int, data))
        elif isinstance(data, list):
            data = array.array("B", data)
        dtype = dtype or self.get_dtype()
        a = self.carrier_amplitude * IQArray.min_max_for_dtype(dtype)[1]
        parameters = self.parameters
        if self.modulation_type == "ASK":
            parameters = array.array("f", [a * p / 100 for p in parameters])
        elif self.modulation_type == "PSK":
            parameters = array.array("f", [p * (math.pi / 180) for p in parameters])
        return data, dtype, a, parameters

    def modulate(self, data=None, pause=0, start=0, dtype=None) -> IQArray:
        assert pause >= 0
        if data is None:
            data = self.data
        else:
            self.data = data
        data, dtype, a, parameters = self._prepared(data, dtype)
        if len(data) == 0:
            return IQArray(None, np.float32, 0)
        result = signal_functions.modulate_c(
            data, self.samples_per_symbol, self.modulation_type, parameters, self.bits_per_symbol, a, self.carrier_freq_hz,
            self.carrier_phase_deg * (np.pi / 180), self.sample_rate, pause, start, dtype, self.gauss_bt, self.gauss_filter_width)
        return IQArray(result, _owned=True)

    def modulate_batch(self, messages, pauses, start=0, dtype=None) -> list:
        """All messages in one GPU batch (what modulate_messages / ContinuousModulator loop over in the reference);
        element m equals modulate(messages[m], pauses[m], start).data"""
        prepared = [self._prepared(m, dtype) for m in messages]
        if not prepared:
            return []
        _, dt, a, parameters = prepared[0]
        res = signal_functions.modulate_batch(
            [p[0] for p in prepared], self.samples_per_symbol, self.modulation_type, parameters, self.bits_per_symbol, a,
            self.carrier_freq_hz, self.carrier_phase_deg * (np.pi / 180), self.sample_rate, pauses, start, dt, self.gauss_bt,
            self.gauss_filter_width)
        return [IQArray(r, _owned=True) for r in res]

    def get_default_parameters(self) -> array.array:
        if self.is_amplitude_based:
            parameters = np.linspace(0, 100, self.modulation_order, dtype=np.float32)
        elif self.is_frequency_based:
            parameters = [(i + 1) * self.carrier_freq_hz / self.modulation_order for i in range(self.modulation_order)]
        elif self.is_phase_based:
            step = 360 / self.modulation_order
            parameters = np.arange(step / 2, 360, step) - 180
            if self.modulation_type == "OQPSK":
                parameters = parameters[[i ^ (i >> 1) for i in range(self.modulation_order)]]  # Gray code order
        else:
            return None
        return array.array("f", parameters)
