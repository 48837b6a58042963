"""CPU: the restatement of numpy's float32 pairwise summation / np.var that pairwise.cu replays on the device equals numpy bit for
bit (the summation tree depends on n only: 8 accumulators per <=128-element leaf, halves split at multiples of 8)."""
import numpy as np


def test_pairwise_sum_and_var_models_equal_numpy(oracle):
    rng = np.random.default_rng(0)
    sizes = list(range(1, 300)) + [1000, 1023, 1024, 1025, 4097, 65537, 100003]
    for n in sizes:
        a = (rng.standard_normal(n) * rng.choice([1, 1e3, 1e-3]) + rng.choice([0, 0.3, -3])).astype(np.float32)
        s = np.float32(0.0) + oracle.np_pairwise_sum_f32(a)
        assert np.float32(s).view(np.uint32) == np.float32(np.add.reduce(a)).view(np.uint32), n
        if n in (1, 2, 7, 8, 9, 127, 128, 129, 255, 1000, 4097, 65537, 100003):
            mean, var = oracle.np_var_f32(a)
            assert np.float32(var).view(np.uint32) == np.float32(np.var(a)).view(np.uint32), n
            assert np.float32(mean).view(np.uint32) == np.float32(np.mean(a)).view(np.uint32), n
