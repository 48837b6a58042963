"""GPU: ContinuousModulator — shaped like the reference's tests/test_continuous_modulator.py:15 (a real child process
modulates a message list into the shared ring buffer), plus what the reference's test does not check: the samples that
arrive in the ring buffer are exactly Modulator.modulate() of the messages, in order."""
import time
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NUM_MESSAGES = 20
BITS_PER_MESSAGE = 100


def _messages():
    # what ContinuousModulator reads of a urh Message: encoded_bits, pause, modulator_index (Message.py is out of scope)
    return [SimpleNamespace(encoded_bits=[True] * BITS_PER_MESSAGE, pause=1000, modulator_index=0) for _ in range(NUM_MESSAGES)]


def _wait(predicate, seconds):
    deadline = time.time() + seconds
    while time.time() < deadline:
        if predicate():
            return True
        time.sleep(0.05)
    return predicate()


def test_modulate_continuously():
    from urh_b200.signalprocessing.ContinuousModulator import ContinuousModulator
    from urh_b200.signalprocessing.Modulator import Modulator

    modulator = Modulator("Test")
    cm = ContinuousModulator(_messages(), [modulator])
    assert cm.current_message_index.value == 0
    assert cm.ring_buffer.is_empty
    cm.start()
    try:
        assert cm.process.is_alive()
        # the child is spawned (CUDA cannot be forked): interpreter start + library load + context creation on a cold box
        assert _wait(lambda: not cm.ring_buffer.is_empty, 120), "nothing arrived in the ring buffer"
        expected = modulator.modulate([True] * BITS_PER_MESSAGE, pause=1000).data
        assert _wait(lambda: len(cm.ring_buffer) >= 2 * len(expected), 60)
        got = cm.ring_buffer.pop(2 * len(expected))
        assert got.dtype == expected.dtype
        assert np.array_equal(got[:len(expected)].view(np.uint32), expected.view(np.uint32))   # message 0
        assert np.array_equal(got[len(expected):].view(np.uint32), expected.view(np.uint32))   # message 1
    finally:
        cm.stop()
    assert not cm.process.is_alive()


def test_finite_repeats_and_mixed_modulators():
    """num_repeats rounds over messages that alternate between two modulators (batches split at modulator changes)"""
    from urh_b200.signalprocessing.ContinuousModulator import ContinuousModulator
    from urh_b200.signalprocessing.Modulator import Modulator

    fsk = Modulator("fsk")
    fsk.modulation_type = "FSK"
    ask = Modulator("ask")
    ask.modulation_type = "ASK"
    rng = np.random.default_rng(4)
    msgs = []
    for i in range(7):
        bits = [bool(b) for b in rng.integers(0, 2, 16 + i)]
        msgs.append(SimpleNamespace(encoded_bits=bits, pause=50 * i, modulator_index=(i // 2) % 2))
    cm = ContinuousModulator(msgs, [fsk, ask], num_repeats=2)
    cm.start()
    try:
        want = np.concatenate([[fsk, ask][m.modulator_index].modulate(m.encoded_bits, pause=m.pause).data for m in msgs] * 2)
        assert _wait(lambda: len(cm.ring_buffer) >= len(want), 120)
        assert _wait(lambda: not cm.process.is_alive(), 30)   # two rounds, then the child ends by itself
        got = cm.ring_buffer.pop(len(want))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert cm.ring_buffer.is_empty
    finally:
        cm.stop()
