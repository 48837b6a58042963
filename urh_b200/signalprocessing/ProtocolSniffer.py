"""``ProtocolSniffer`` — the live demodulation loop as a streaming client of the GPU path (reference:
src/urh/signalprocessing/ProtocolSniffer.py:20-283).

Only the signal path: ``feed(data)`` is the reference's ``__demodulate_data`` (ProtocolSniffer.py:204-283) — noise gate on the
chunk's RMS power, adaptive noise threshold, buffering until 10 symbol lengths of silence (or a full buffer) end a message, then
Signal.qad -> detect_center (optional) -> grab_pulse_lens -> pulses -> bits on the device.  The SDR backends, the receive thread and
the Qt signals around it (VirtualDevice, check_for_data, sniff files) are out of scope: a caller pushes chunks of IQ samples."""
import time

import numpy as np

from ..ainterpretation import AutoInterpretation
from ..cythonext.signal_functions import grab_pulse_lens
from .IQArray import IQArray
from .ProtocolAnalyzer import LiteMessage, ProtocolAnalyzer
from .Signal import Signal


class ProtocolSniffer(ProtocolAnalyzer):
    BUFFER_SIZE_MB = 100

    def __init__(self, samples_per_symbol: int, center: float, center_spacing: float, noise: float, tolerance: int, modulation_type: str,
                 bits_per_symbol: int, sample_rate: float = 1e6, data_type=np.float32):
        signal = Signal("", "LiveSignal")
        signal.samples_per_symbol = samples_per_symbol
        signal.center = center
        signal.center_spacing = center_spacing
        signal.noise_threshold = noise
        signal.tolerance = tolerance
        signal.silent_set_modulation_type(modulation_type)
        signal.bits_per_symbol = bits_per_symbol
        ProtocolAnalyzer.__init__(self, signal)
        self.sample_rate = sample_rate
        self.data_type = data_type
        signal.iq_array = IQArray(None, data_type, 0)
        self.adaptive_noise = False
        self.automatic_center = False
        self.pause_length = 0
        self.__init_buffer()

    # -- buffer (ProtocolSniffer.py:86-108) ---------------------------------------------------------------------------------------
    def __init_buffer(self):
        self.__buffer = IQArray(None, self.data_type, int(self.BUFFER_SIZE_MB * 1000 * 1000 / 8))
        self.__current_buffer_index = 0

    def __add_to_buffer(self, data: np.ndarray):
        n = len(data)
        if n + self.__current_buffer_index > len(self.__buffer):
            n = len(self.__buffer) - self.__current_buffer_index - 1
        self.__buffer[self.__current_buffer_index: self.__current_buffer_index + n] = data[:n]
        self.__current_buffer_index += n

    def __clear_buffer(self):
        self.__current_buffer_index = 0

    def __buffer_is_full(self):
        return self.__current_buffer_index >= len(self.__buffer) - 2

    def clear(self):
        self.__clear_buffer()
        self.messages.clear()

    # -- ProtocolSniffer.py:204-283 ------------------------------------------------------------------------------------------------------
    def feed(self, data: np.ndarray):
        """one chunk of received samples (complex64, as the reference's raw-mode receive buffer delivers them)"""
        if len(data) == 0:
            return
        power_spectrum = data.real ** 2.0 + data.imag ** 2.0
        is_above_noise = np.sqrt(np.mean(power_spectrum)) > self.signal.noise_threshold
        if self.adaptive_noise and not is_above_noise:
            self.signal.noise_threshold = 0.9 * self.signal.noise_threshold + 0.1 * np.sqrt(np.max(power_spectrum))
        if is_above_noise:
            self.__add_to_buffer(data)
            self.pause_length = 0
            if not self.__buffer_is_full():
                return
        else:
            self.pause_length += len(data)
            if self.pause_length < 10 * self.signal.samples_per_symbol:
                self.__add_to_buffer(data)
                if not self.__buffer_is_full():
                    return
        if self.__current_buffer_index == 0:
            return
        self.flush()

    def flush(self):
        """demodulate what the buffer holds and append the messages found (the tail of __demodulate_data)"""
        if self.__current_buffer_index == 0:
            return
        self.signal.iq_array = IQArray(np.array(self.__buffer._peek(slice(0, self.__current_buffer_index)), order="C"), _owned=True)
        self.signal.timestamp = time.time() - (len(self.signal.iq_array) / self.sample_rate)
        self.__clear_buffer()
        self.signal._qad = None
        self.signal._qad_dev = None
        samples_per_symbol = self.signal.samples_per_symbol
        qad = self.signal.qad_device
        if self.automatic_center:
            self.signal.center = AutoInterpretation.detect_center(qad, max_size=150 * samples_per_symbol)
        ppseq = grab_pulse_lens(qad, self.signal.center, self.signal.tolerance, self.signal.modulation_type, self.signal.samples_per_symbol,
                                self.signal.bits_per_symbol, self.signal.center_spacing)
        bit_data, pauses, bit_sample_pos = self._ppseq_to_bits_device(ppseq, samples_per_symbol, self.signal.bits_per_symbol,
                                                                      write_bit_sample_pos=True)
        for i, (bits, pause) in enumerate(zip(bit_data, pauses)):
            message = LiteMessage(bits, pause, bit_sample_pos[i])
            message.timestamp = self.signal.timestamp + (bit_sample_pos[i][0] / self.sample_rate)
            self.messages.append(message)
