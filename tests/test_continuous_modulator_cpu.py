"""CPU: the process / ring-buffer mechanics of ContinuousModulator (reference: tests/test_continuous_modulator.py:15) with a stub
in place of the GPU modulator — a real spawned child, batches split at modulator changes, finite repeats, stop()."""
import time
from types import SimpleNamespace

import numpy as np


class StubModulator(object):
    """modulate_batch of a message = its bits as (n, 2) float32 samples, followed by `pause` zero samples"""

    def __init__(self, scale):
        self.scale = scale

    def modulate_batch(self, messages, pauses, start=0, dtype=None):
        out = []
        for bits, pause in zip(messages, pauses):
            a = np.zeros((len(bits) + pause, 2), dtype=np.float32)
            a[:len(bits), 0] = np.asarray(bits, dtype=np.float32) * self.scale
            a[:len(bits), 1] = np.arange(len(bits), dtype=np.float32)
            out.append(SimpleNamespace(data=a, __len__=None))
        return [_Arr(o.data) for o in out]


class _Arr(object):
    """what ContinuousModulator needs of an IQArray: len() and something RingBuffer.push can index"""

    def __init__(self, a):
        self.a = a

    def __len__(self):
        return len(self.a)

    def __getitem__(self, item):
        return self.a[item]


def _wait(predicate, seconds):
    deadline = time.time() + seconds
    while time.time() < deadline:
        if predicate():
            return True
        time.sleep(0.02)
    return predicate()


def _messages(n, rng):
    return [SimpleNamespace(encoded_bits=[int(b) for b in rng.integers(0, 2, 8 + i)], pause=3 * i, modulator_index=(i // 3) % 2)
            for i in range(n)]


def test_child_fills_ring_buffer_in_order():
    from urh_b200.signalprocessing.ContinuousModulator import ContinuousModulator

    rng = np.random.default_rng(0)
    msgs = _messages(10, rng)
    mods = [StubModulator(1.0), StubModulator(-2.0)]
    cm = ContinuousModulator(msgs, mods, num_repeats=2)
    assert cm.current_message_index.value == 0 and cm.ring_buffer.is_empty
    cm.start()
    try:
        assert cm.process.is_alive() or cm.process.exitcode == 0
        want = np.concatenate([mods[m.modulator_index].modulate_batch([m.encoded_bits], [m.pause])[0].a for m in msgs] * 2)
        assert _wait(lambda: len(cm.ring_buffer) >= len(want), 60)
        assert _wait(lambda: not cm.process.is_alive(), 30)
        assert cm.process.exitcode == 0
        got = cm.ring_buffer.pop(len(want))
        assert np.array_equal(got, want)
    finally:
        cm.stop()
    assert not cm.process.is_alive()


def test_infinite_mode_stops_on_request():
    from urh_b200.signalprocessing.ContinuousModulator import ContinuousModulator

    msgs = _messages(4, np.random.default_rng(1))
    cm = ContinuousModulator(msgs, [StubModulator(1.0), StubModulator(1.0)])   # num_repeats = -1: forever
    cm.start()
    assert _wait(lambda: not cm.ring_buffer.is_empty, 60)
    assert cm.process.is_alive()
    cm.stop()
    assert not cm.process.is_alive()
    assert cm.ring_buffer.left_index == 0 and cm.ring_buffer.right_index == 0   # clear() as in the reference (RingBuffer.py:73-75)
