"""CPU model of the whole-tile boundary resolve (urh_b200/csrc/dense.cuh: UrhTileResolve) against the serial run tracker it replaced.

A tile is 32 groups of 64 samples; lane g keeps group g's class masks (bit l of n0/a0: sample 2l is noise / above the threshold,
n1/a1: sample 2l+1) and settles its own boundaries after one prefix-max ("where did the run entering my group start") and one
prefix-sum (slots of the candidates) over the lanes.  This file restates both forms in Python, mask arithmetic included, and checks
that they produce the same tile summary and the same staged candidates for random class sequences, every tolerance regime and the
edge cases (one run, a boundary at every sample, boundaries on group and lane edges).  The CUDA kernel itself is checked against
the oracle by the -m gpu tests; this pins the restatement the kernel implements."""
import numpy as np
import pytest

TILE = 2048


def serial_tracker(cls, tol):
    """UrhRunTracker over one full tile: (first_cls, last_cls, head_len, tail_len, [(pos, cls) ...])"""
    n = len(cls)
    run_start, run_cls, is_head, head_len, cands = 0, int(cls[0]), True, n, []
    for p in range(1, n):
        if cls[p] == cls[p - 1]:
            continue
        if is_head:
            head_len, is_head = p, False
        elif p - run_start > tol:
            cands.append((run_start + tol, run_cls))
        run_start, run_cls = p, int(cls[p])
    if not is_head and n - run_start > tol:
        cands.append((run_start + tol, run_cls))
    return int(cls[0]), run_cls, head_len, n - run_start, cands


def masks_of(cls):
    """per group g: (n0, a0, n1, a1) as 32-bit integers"""
    out = []
    for g in range(TILE // 64):
        c = cls[g * 64:(g + 1) * 64]
        n0 = a0 = n1 = a1 = 0
        for l in range(32):
            if c[2 * l] < 0:
                n0 |= 1 << l
            elif c[2 * l] == 1:
                a0 |= 1 << l
            if c[2 * l + 1] < 0:
                n1 |= 1 << l
            elif c[2 * l + 1] == 1:
                a1 |= 1 << l
        out.append((n0, a0, n1, a1))
    return out


def cls_of(nbit, abit):
    return -1 if (nbit & 1) else (abit & 1)


def ffs(x):
    return (x & -x).bit_length()   # 1-based, 0 for x == 0


def clz32(x):
    return 32 - x.bit_length()


def lane_walk(lane, m0, m1, masks, pn, pa, q, tol):
    """UrhTileResolve::walk: the candidates of lane's group, q = start of the run entering it"""
    n0, a0, n1, a1 = masks
    out = []
    while m0 | m1:
        l0 = ffs(m0) - 1 if m0 else 64
        l1 = ffs(m1) - 1 if m1 else 64
        take0 = l0 <= l1
        l = l0 if take0 else l1
        if take0:
            m0 &= m0 - 1
        else:
            m1 &= m1 - 1
        p = lane * 64 + 2 * l + (0 if take0 else 1)
        if q > 0 and p - q > tol:
            if not take0:
                nb, ab = n0 >> l, a0 >> l
            elif l > 0:
                nb, ab = n1 >> (l - 1), a1 >> (l - 1)
            else:
                nb, ab = pn, pa
            out.append((q + tol, cls_of(nb, ab)))
        q = p
    return out


def tile_resolve(cls, tol):
    """UrhTileResolve::finish, lane by lane"""
    M = masks_of(cls)
    lanes = len(M)
    m0s, m1s, last, pns, pas = [], [], [], [], []
    for g, (n0, a0, n1, a1) in enumerate(M):
        pn = (M[g - 1][2] >> 31) & 1 if g else 0
        pa = (M[g - 1][3] >> 31) & 1 if g else 0
        m0 = ((((n1 << 1) & 0xffffffff) | pn) ^ n0) | ((((a1 << 1) & 0xffffffff) | pa) ^ a0)
        if g == 0:
            m0 |= 1
        m1 = (n0 ^ n1) | (a0 ^ a1)
        lp = -1
        if m1:
            lp = g * 64 + 2 * (31 - clz32(m1)) + 1
        if m0:
            lp = max(lp, g * 64 + 2 * (31 - clz32(m0)))
        m0s.append(m0); m1s.append(m1); last.append(lp); pns.append(pn); pas.append(pa)
    incl = np.maximum.accumulate(np.array(last))
    q = [-1] + [int(x) for x in incl[:-1]]
    L = int(incl[-1])
    first = TILE
    for g in range(lanes):
        f0, f1 = m0s[g] & (~1 if g == 0 else 0xffffffff), m1s[g]
        if f0 | f1:
            l0 = ffs(f0) - 1 if f0 else 64
            l1 = ffs(f1) - 1 if f1 else 64
            first = min(first, g * 64 + (2 * l0 if l0 <= l1 else 2 * l1 + 1))
    cands = []
    for g in range(lanes):   # the prefix sum of the counts is the concatenation order
        cands += lane_walk(g, m0s[g], m1s[g], M[g], pns[g], pas[g], q[g], tol)
    last_cls = cls_of(M[-1][2] >> 31, M[-1][3] >> 31)
    if L > 0 and TILE - L > tol:
        cands.append((L + tol, last_cls))
    return cls_of(M[0][0], M[0][1]), last_cls, first, TILE - L, cands


def random_classes(rng, mean_run, noise_share):
    out = np.empty(TILE, dtype=np.int64)
    p = 0
    prev = None
    while p < TILE:
        r = 1 + int(rng.geometric(1.0 / mean_run)) if mean_run > 1 else 1
        c = -1 if rng.random() < noise_share else int(rng.integers(0, 2))
        if c == prev:
            c = 1 - c if c >= 0 else int(rng.integers(0, 2))
        out[p:p + r] = c
        prev = c
        p += r
    return out


@pytest.mark.parametrize("tol", [0, 1, 5, 63, 64, 100, 2047, 5000])
@pytest.mark.parametrize("mean_run", [1, 3, 40, 100, 700])
def test_resolve_equals_serial_tracker(tol, mean_run):
    rng = np.random.default_rng(1000 * tol + mean_run)
    for trial in range(6):
        cls = random_classes(rng, mean_run, noise_share=0.0 if trial % 2 else 0.2)
        assert tile_resolve(cls, tol) == serial_tracker(cls, tol)


@pytest.mark.parametrize("tol", [0, 5, 64])
def test_resolve_edge_cases(tol):
    cases = []
    for c in (-1, 0, 1):
        cases.append(np.full(TILE, c))                                   # one run: the head run is the whole tile
    alt = np.arange(TILE) % 2
    cases.append(alt)                                                    # a boundary at every sample
    cases.append(np.where(np.arange(TILE) % 2 == 0, -1, 1))              # ... between noise and a class
    for edge in (1, 2, 63, 64, 65, 127, 128, 1024, 2046, 2047):          # one boundary on lane / group edges
        c = np.zeros(TILE, dtype=np.int64)
        c[edge:] = 1
        cases.append(c)
        c2 = c.copy()
        c2[min(edge + tol + 1, TILE - 1):] = -1                          # a second one right after the tolerance
        cases.append(c2)
    for cls in cases:
        assert tile_resolve(np.asarray(cls, dtype=np.int64), tol) == serial_tracker(np.asarray(cls, dtype=np.int64), tol)
