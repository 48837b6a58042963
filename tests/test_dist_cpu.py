"""CPU suite (gloo, world_size 2): host-side logic of the sharded-capture protocol (urh_b200/dist.py) — the
run-carry fold, the shard bounds and the exchange helpers — checked end to end against the oracle's serial
digitizer with a numpy stand-in for the per-rank dense pass."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def np_classes(x, center, noise_value):
    return np.where(x == noise_value, -1, np.where(x <= center, 0, 1)).astype(np.int64)


def np_shard_summary(cls):
    """(last_cls, last_len, whole) of a shard's class sequence"""
    change = np.nonzero(np.diff(cls))[0]
    start = 0 if len(change) == 0 else int(change[-1]) + 1
    return int(cls[-1]), len(cls) - start, len(change) == 0


def np_shard_candidates(cls, tol, carry, offset):
    """candidates (global pos, class) of runs longer than tol; the first run continues `carry` = (cls, len) or None"""
    bounds = np.concatenate(([0], np.nonzero(np.diff(cls))[0] + 1, [len(cls)]))
    out = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        c = int(cls[a])
        before = carry[1] if (a == 0 and carry is not None and carry[0] == c) else 0
        if before <= tol < before + (b - a):
            out.append((offset + a + tol - before, c))
    return out


def np_pulses(cands, n, tol, is_ask, sps, init):
    rows = []
    prev_cls, prev_pos = init, None
    for pos, c in cands:
        if c == prev_cls:
            continue
        rec = pos + 1 - tol if prev_pos is None else pos - prev_pos
        state = prev_cls
        if is_ask and state == -1 and rec < sps:
            state = 0
        if rows and rows[-1][0] == state:
            rows[-1][1] += rec
        else:
            rows.append([state, rec])
        prev_cls, prev_pos = c, pos
    rec = n - tol if prev_pos is None else n - 1 - prev_pos
    if len(rows) < n:
        if rows and rows[-1][0] == prev_cls:
            rows[-1][1] += rec
        else:
            rows.append([prev_cls, rec])
    return np.array(rows, dtype=np.int64).reshape(-1, 2)


def test_fold_carry_and_bounds():
    from urh_b200.dist import fold_carry, shard_bounds

    assert fold_carry([(1, 5, False)]) == [None]
    assert fold_carry([(1, 5, False), (1, 10, True), (1, 3, False), (0, 2, False)]) == [None, (1, 5), (1, 15), (1, 3)]
    assert fold_carry([(0, 8, True), (0, 8, True), (1, 1, False)]) == [None, (0, 8), (0, 16)]
    assert fold_carry([(0, 8, True), (1, 8, True), (1, 4, True)]) == [None, (0, 8), (1, 8)]
    b = shard_bounds(10_000_000, 4)
    assert b[0][0] == 0 and b[-1][1] == 10_000_000 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert all(s % 2048 == 0 for s, _ in b)
    assert shard_bounds(1000, 8)[0] == (0, 1000)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    from urh_b200.dist import HostExchange, fold_carry, shard_bounds
    from oracle import oracle

    hx = HostExchange()
    rng = np.random.default_rng(123)  # same data on every rank
    ok = True
    for trial in range(25):
        n = int(rng.integers(50, 9000))
        period = int(rng.integers(2, 80))
        x = (np.repeat(rng.standard_normal(n // period + 1), period)[:n] * 0.5 + 0.1 * rng.standard_normal(n)).astype(np.float32)
        x[rng.random(n) < 0.05] = -4.0
        s = int(rng.integers(0, n))
        x[s: s + int(rng.integers(0, 4000))] = -4.0
        tol = int(rng.choice([0, 1, 5, 40, 3000]))
        bounds = shard_bounds(n, world, align=int(rng.choice([1, 7, 64])))
        lo, hi = bounds[rank]
        cls = np_classes(x[lo:hi], 0.05, -4.0) if hi > lo else np.zeros(0, np.int64)
        if hi > lo:
            summary = np_shard_summary(cls)
        else:
            summary = None
        every = hx.allgather(summary)
        # empty shards cannot occur with shard_bounds (the last rank takes the rest) except when n < world
        live = [e for e in every if e is not None]
        carries = fold_carry(live)
        idx = sum(1 for e in every[:rank] if e is not None)
        cands = np_shard_candidates(cls, tol, carries[idx], lo) if hi > lo else []
        gathered = hx.allgather(cands)
        if rank == 0:
            allc = [c for part in gathered for c in part]
            init = -1 if x[0] == -4.0 else (0 if 0.0 <= 0.05 else 1)
            rows = np_pulses(allc, n, tol, False, 20, init)
            ref = oracle.grab_pulse_lens(x, 0.05, tol, "FSK", 20)
            ok = ok and np.array_equal(rows, ref)
    res = hx.allgather(bool(ok))
    if rank == 0:
        open(os.path.join(tmp, "ok"), "w").write("1" if all(res) else "0")
    dist.destroy_process_group()


def test_sharded_protocol_matches_serial_digitizer_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"


# ---- capture-wide detect_center over shards: the exchange protocol of urh_b200.dist.center_protocol over gloo -------------
def _center_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    from urh_b200.dist import HostExchange, center_protocol, shard_bounds
    from oracle import oracle

    hx = HostExchange()
    rng = np.random.default_rng(77)  # same data on every rank
    ok = True
    for trial in range(20):
        n = int(rng.integers(200, 30000))
        levels = rng.choice([-0.3, 0.3], n // 50 + 1) if trial % 3 else rng.choice([-1.0, -0.2, 0.2, 1.0], n // 50 + 1)
        x = (np.repeat(levels, 50)[:n] + 0.02 * rng.standard_normal(n)).astype(np.float32)
        x[rng.random(n) < 0.1] = -4.0
        if trial % 4 == 0:
            s = int(rng.integers(0, n))
            x[s: s + n // 3] = -4.0                      # a long gap: a shard may keep nothing at all
        if trial == 7:
            x[:] = -4.0                                   # nothing kept anywhere -> None
        if trial == 9:
            x[x > -4] = 0.25                              # constant signal: zero variance -> None
        max_size = None if trial % 5 else 500
        lo, hi = shard_bounds(n, world, align=int(rng.choice([1, 64])))[rank]
        mine = x[lo:hi]
        kept = mine[mine > -4]

        def window_stats(a, b):
            r = kept[a:b].astype(np.float64)
            if len(r) == 0:
                return [0.0, np.inf, -np.inf, 0.0, 0.0]
            return [len(r), r.min(), r.max(), r.sum(), (r * r).sum()]

        def histogram(a, b, edges):
            exact = edges[0] + np.arange(len(edges)) * (edges[1] - edges[0])
            return np.histogram(kept[a:b], bins=exact)[0]

        def allgather(v):
            return np.array(hx.allgather(np.asarray(v, dtype=np.int64)))

        def allreduce(y):
            t = torch.from_numpy(np.ascontiguousarray(y))
            dist.all_reduce(t)
            return t.numpy()

        c = center_protocol(rank, world, len(kept), window_stats, histogram, allgather, allreduce, max_size)
        ref = oracle.detect_center(x, max_size=max_size)
        same = (c is None) == (ref is None) and (c is None or abs(c - ref) <= 2e-6 * max(1.0, abs(ref)))
        ok = ok and same
    res = hx.allgather(bool(ok))
    if rank == 0:
        open(os.path.join(tmp, "ok_center"), "w").write("1" if all(res) else "0")
    dist.destroy_process_group()


def test_distributed_center_protocol_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = 29450 + os.getpid() % 200
    mp.spawn(_center_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok_center").read() == "1"


def test_resolve_psk_chain():
    """host half of the sharded Costas loop: each shard adopts the hypothesis whose start state is the predecessor's end state"""
    import numpy as np
    from urh_b200.dist import resolve_psk_chain

    k = lambda a, b: np.array([a, b], np.float32).tobytes()  # noqa: E731
    hyps = [[(k(0, 1.5), k(1, 2))],
            [(k(9, 9), k(3, 3)), (k(1, 2), k(4, 4))],
            [(k(4, 4), k(5, 5)), (k(0, 0), k(6, 6))]]
    assert resolve_psk_chain(hyps) == ([0, 1, 0], 3)
    hyps[2][0] = (k(7, 7), k(5, 5))               # shard 2 starts in no hypothesis' state: hand-over from shard 2 on
    assert resolve_psk_chain(hyps) == ([0, 1], 2)
    hyps[1] = [(k(-0.0, 2), k(3, 3))]             # -0.0 != +0.0 bitwise... and 1 != -0: unresolved from shard 1
    assert resolve_psk_chain(hyps) == ([0], 1)
    assert resolve_psk_chain(hyps[:1]) == ([0], 1)
