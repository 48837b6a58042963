"""Shared-memory ring buffer of IQ samples between the continuous modulator process and its consumer
(reference: src/urh/util/RingBuffer.py:7-140).  Host plumbing around the modulator; same interface."""
from multiprocessing import get_context

import numpy as np

_TYPECODES = {np.uint8: "B", np.int8: "b", np.int16: "h", np.uint16: "H", np.float32: "f", np.float64: "d"}


class RingBuffer(object):
    def __init__(self, size: int, dtype=np.float32):
        self.dtype = dtype
        self.size = size
        # shared objects from the SPAWN context: the producer is a spawned child (CUDA cannot be forked, and URH itself
        # forces the spawn start method); spawn-context locks are also fine under fork
        mp = get_context("spawn")
        self.__data = mp.Array(_TYPECODES[dtype], 2 * size)
        self.__left = mp.Value("L", 0)
        self.__right = mp.Value("L", 0)
        self.__length = mp.Value("L", 0)

    def __len__(self):
        return self.__length.value

    @property
    def left_index(self):
        return self.__left.value

    @left_index.setter
    def left_index(self, value):
        self.__left.value = value % self.size

    @property
    def right_index(self):
        return self.__right.value

    @right_index.setter
    def right_index(self, value):
        self.__right.value = value % self.size

    @property
    def is_empty(self) -> bool:
        return len(self) == 0

    @property
    def space_left(self):
        return self.size - len(self)

    def _view(self):
        return np.frombuffer(self.__data.get_obj(), dtype=self.dtype).reshape(len(self.__data) // 2, 2)

    @property
    def data(self):
        return self._view()

    @property
    def view_data(self):
        left, right = self.left_index, self.left_index + len(self)
        if left > right:
            left, right = right, left
        flat = self.data.flatten()
        return np.concatenate((flat[left:right], flat[right:], flat[:left]))

    def clear(self):
        self.left_index = 0
        self.right_index = 0

    def will_fit(self, number_values: int) -> bool:
        return number_values <= self.space_left

    def push(self, values):
        n = len(values)
        if len(self) + n > self.size:
            raise ValueError("Too much data to push to RingBuffer")
        head = min(n, self.size - self.right_index)  # samples that fit before the wrap
        with self.__data.get_lock():
            data = self._view()
            data[self.right_index: self.right_index + head] = values[:head]
            data[: n - head] = values[head:]
            self.right_index += n
        self.__length.value += n

    def pop(self, number: int, ensure_even_length=False) -> np.ndarray:
        if ensure_even_length:
            number -= number % 2
        if len(self) == 0 or number == 0:
            return np.array([], dtype=self.dtype)
        number = len(self) if number < 0 else min(number, len(self))
        with self.__data.get_lock():
            data = self._view()
            result = np.ones((number, 2), dtype=self.dtype)
            head = min(number, len(data) - self.left_index)
            result[:head] = data[self.left_index: self.left_index + head]
            if head < number:
                result[head:] = data[: number - head]
        self.left_index += number
        self.__length.value -= number
        return result
