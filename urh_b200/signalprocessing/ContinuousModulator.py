"""``ContinuousModulator`` (reference: src/urh/signalprocessing/ContinuousModulator.py:10-99): a background process
modulates a message list into a shared ring buffer.  Same interface; the child (``spawn``, which CUDA needs and URH
already forces) modulates runs of messages that share a modulator as one GPU batch instead of one call per message."""
import time
from multiprocessing import get_context

from .. import settings
from ..util.RingBuffer import RingBuffer
from .Modulator import Modulator


class ContinuousModulator(object):
    WAIT_TIMEOUT = 0.1
    BATCH = 256  # messages per GPU batch

    def __init__(self, messages, modulators, num_repeats=-1):
        self.messages = messages
        self.modulators = modulators
        self.num_repeats = num_repeats  # -1 or 0 = infinite
        self.ring_buffer = RingBuffer(int(settings.CONTINUOUS_BUFFER_SIZE_MB * 1e6) // 8, dtype=Modulator.get_dtype())
        self._mp = get_context("spawn")
        self.current_message_index = self._mp.Value("L", 0)
        self.abort = self._mp.Value("i", 0)
        self.process = self._mp.Process(target=self.modulate_continuously, args=(self.num_repeats,), daemon=True)

    @property
    def is_running(self):
        return self.process.is_alive()

    def start(self):
        self.abort.value = 0
        self.process = self._mp.Process(target=self.modulate_continuously, args=(self.num_repeats,), daemon=True)
        self.process.start()

    def stop(self, clear_buffer=True):
        self.abort.value = 1
        if self.process.is_alive():
            self.process.join(1.5)
            if self.process.is_alive():
                self.process.terminate()
        if clear_buffer:
            self.ring_buffer.clear()

    def _push(self, modulated) -> bool:
        while not self.ring_buffer.will_fit(len(modulated)):
            if self.abort.value:
                return False
            time.sleep(self.WAIT_TIMEOUT)  # wait for space
        self.ring_buffer.push(modulated)
        return True

    def modulate_continuously(self, num_repeats):
        rounds = iter(int, 1) if num_repeats <= 0 else range(0, num_repeats)  # <= 0: forever
        for _ in rounds:
            if self.abort.value:
                return
            i = self.current_message_index.value
            while i < len(self.messages):
                if self.abort.value:
                    return
                # run of consecutive messages with the same modulator -> one batch
                mod_index = self.messages[i].modulator_index
                j = i
                while j < len(self.messages) and j - i < self.BATCH and self.messages[j].modulator_index == mod_index:
                    j += 1
                modulator = self.modulators[mod_index]
                batch = modulator.modulate_batch([m.encoded_bits for m in self.messages[i:j]],
                                                 [m.pause for m in self.messages[i:j]], start=0)
                for k, modulated in enumerate(batch):
                    self.current_message_index.value = i + k
                    if not self._push(modulated):
                        return
                i = j
            self.current_message_index.value = 0
