"""One capture sharded by contiguous sample range over the GPUs of one box.

One process per GPU (launched by torchrun).  ``torch.distributed`` (gloo) is the launcher plumbing: rendezvous and
the exchange of tiny host-side descriptors; the data path between GPUs is NCCL over NVLink, driven from
liburh_b200 (nccl.cu): the 1-sample halo, the all-reduce of the global noise statistics and the gather of the
sparse candidate tables.  Per rank the sample-rate work is exactly the single-GPU dense pass.

Protocol of the sharded digitizer (exactness argument in DESIGN.md §6):
  1. every rank: dense pass over its shard (+1 halo sample for the FSK conjugate product) -> tile table and the
     summary (class, length, whole?) of the run that closes the shard;
  2. all-gather the summaries; rank r folds those of ranks < r with the run-carry operator -> the run that ends
     right before its shard;
  3. every rank: candidate table with global positions (the carry only affects the first run of the shard);
  4. gather the candidate tables on rank 0 (NCCL send/recv), which finishes exactly like the single-GPU path.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .device import DeviceArray


# ---- pure host logic (unit-tested on CPU with gloo, tests/test_dist_cpu.py) -----------------------------------------
def fold_carry(summaries):
    """summaries: list of (last_cls, last_len, whole) per rank, in rank order.
    Returns for every rank the run that ends right before its shard: None for rank 0, else (cls, len)."""
    out = []
    carry = None  # (cls, len, whole)
    for cls, length, whole in summaries:
        out.append(None if carry is None else (carry[0], carry[1]))
        if carry is not None and whole and cls == carry[0]:
            carry = (cls, carry[1] + length, carry[2])
        else:
            carry = (cls, length, bool(whole) and carry is None)
    return out


def shard_bounds(n_total: int, world: int, align: int = 2048):
    """contiguous shards whose boundaries are multiples of `align` (the dense pass's tile), last one takes the rest"""
    per = (n_total // world) // align * align
    if per == 0:
        per = n_total
    bounds = []
    start = 0
    for r in range(world):
        end = n_total if r == world - 1 else min(n_total, start + per)
        bounds.append((start, end))
        start = end
    return bounds


def combine_noise_chunks(n_total, chunksize, nchunks, partial_sums, partial_maxs):
    """Global end-aligned noise chunks (AutoInterpretation.py:66-72) from per-rank partial (sum, max) arrays that were
    all-reduced element-wise (sum / max) — identity here; kept for symmetry with the CPU test."""
    return np.asarray(partial_sums, dtype=np.float64), np.asarray(partial_maxs, dtype=np.float64)


class HostExchange(object):
    """tiny host-side collectives over torch.distributed (gloo)"""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    def allgather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def broadcast(self, obj, src=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=src)
        return box[0]

    def barrier(self):
        self.dist.barrier()


def init_nccl(ctx: _lib.Context, hx: HostExchange):
    """create the NCCL communicator of liburh_b200 for this context (id from rank 0 via the host exchange)"""
    buf = C.create_string_buffer(128)
    if hx.rank == 0:
        rc = ctx.lib.urh_nccl_unique_id(buf)
        if rc != 0:
            raise RuntimeError("urh_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    ident = hx.broadcast(bytes(buf.raw) if hx.rank == 0 else None, src=0)
    ctx.check(ctx.lib.urh_nccl_init(ctx.handle, C.c_char_p(ident), hx.rank, hx.world))
    init_p2p(ctx, hx)


def init_p2p(ctx: _lib.Context, hx: HostExchange):
    """NVLink peer mailboxes for the few-bytes exchanges (p2p.cu): host-side all-gathers and the device-resident, stream-ordered
    all-gather / histogram sum the sharded chains use between their kernels.  Used only if EVERY rank could map every peer (one
    node, <= 8 GPUs, CUDA IPC available); otherwise those exchanges stay on NCCL.
    Opt-in with URH_B200_P2P=1.  Measured at 2 GPUs (round 2): 8.3 us per device all-gather against 7.2 us for NCCL's, 9.1 / 14.9 us
    for the histogram sum of 8 / 6000 words against 10.4 us; the step is 4.23 ms with the mailboxes and 4.21 ms with NCCL -- the
    per-step cost of a sharded capture is rank skew and the small device stages between the exchanges, not the collective's
    latency -- so NCCL stays the default."""
    ctx.p2p = False
    ok = hx.world <= 8 and hx.world > 1 and os.environ.get("URH_B200_P2P", "0") == "1"
    handle = C.create_string_buffer(64)
    if ok:
        ok = ctx.lib.urh_p2p_create(ctx.handle, handle) == 0
    handles = hx.allgather(bytes(handle.raw) if ok else None)
    ok = ok and all(h is not None for h in handles)
    if ok:
        ok = ctx.lib.urh_p2p_open(ctx.handle, C.c_char_p(b"".join(handles)), hx.rank, hx.world) == 0
    ctx.p2p = all(hx.allgather(bool(ok)))
    if ok and not ctx.p2p:
        ctx.lib.urh_p2p_close(ctx.handle)


class ShardBuffer(object):
    """Device buffer [pad][halo][shard samples...]: the shard starts 256-byte aligned (full-line warp loads), the halo
    sample sits right before it."""

    def __init__(self, ctx, n_local, dtype=np.float32, halo=1):
        self.ctx = ctx
        self.n = int(n_local)
        self.dtype = np.dtype(dtype)
        self.halo_len = int(halo)
        unit = 256 // (2 * self.dtype.itemsize)  # samples per 256 bytes
        self.pad = ((self.halo_len + unit - 1) // unit) * unit  # samples before the shard (keeps 256 B alignment)
        self.buf = DeviceArray(ctx, (self.n + self.pad, 2), self.dtype)
        self.shard = self.buf[self.pad:]
        self.halo = self.buf[self.pad - self.halo_len: self.pad]


def exchange_halo(ctx, hx, sb: ShardBuffer):
    """rank r receives the last `halo_len` samples of rank r-1's shard (NCCL all-gather of the tails)"""
    h = sb.halo_len
    assert sb.n >= h, "shard shorter than the halo"
    tail = sb.shard[sb.n - h: sb.n]
    allv = DeviceArray(ctx, (hx.world * h, 2), sb.dtype)
    ctx.check(ctx.lib.urh_nccl_allgather(ctx.handle, C.c_void_p(tail.ptr), C.c_void_p(allv.ptr), tail.nbytes))
    if hx.rank > 0:
        src = allv[(hx.rank - 1) * h: hx.rank * h]
        ctx.check(ctx.lib.urh_memcpy_d2d(ctx.handle, C.c_void_p(sb.halo.ptr), C.c_void_p(src.ptr), src.nbytes))
    ctx.sync()


def costas_halo(ctx) -> int:
    return int(ctx.lib.urh_costas_halo_samples())


def resolve_psk_chain(hyps):
    """Pure host logic of the sharded Costas loop (unit-tested on the CPU).  ``hyps[r]`` = list of (start_state, end_state) per
    hypothesis of rank r, states as raw 8-byte keys; rank 0 has one entry (its true run).  Returns (picks, first_unresolved):
    picks[r] = hypothesis of rank r whose start state equals the true end state of rank r-1, for r < first_unresolved;
    first_unresolved = world when every shard is resolved."""
    picks = [0]
    state = hyps[0][0][1]
    for r in range(1, len(hyps)):
        match = [h for h, (start, _) in enumerate(hyps[r]) if start == state]
        if not match:
            return picks, r
        picks.append(match[0])
        state = hyps[r][match[0]][1]
    return picks, len(hyps)


def afp_demod_psk_sharded(ctx, rank, world, sb: ShardBuffer, noise_mag, mod_order, costas_loop_bandwidth, d_out):
    """PSK demodulation (Costas loop) of a capture sharded over the ranks, bit-identical to the serial loop.  Every rank
    speculates over its shard concurrently (the expensive pass) AND hops over it under each hypothesis "my shard starts in
    candidate h's start state" — what a locked loop of the preceding shard ends in, bit for bit.  One all-gather of
    (start, end) state pairs lets every rank pick its hypothesis (``resolve_psk_chain``); nothing waits for a neighbour.
    Only a shard whose predecessor ends in no hypothesis' start state (it begins inside a gap that carries a frozen or creeping
    loop state) falls back to the rank-to-rank hand-over from that shard on.  sb needs a halo of costas_halo() samples."""
    lib = ctx.lib
    assert rank == 0 or sb.halo_len >= costas_halo(ctx)
    ctx.check(lib.urh_costas_shard_speculate(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.dtype_code(sb.dtype), sb.n, int(rank == 0),
                                             float(noise_mag), int(mod_order), float(costas_loop_bandwidth), C.c_void_p(d_out.ptr)))
    mine = np.zeros((4, 4), dtype=np.float32)
    count = C.c_int(0)
    ctx.check(lib.urh_costas_shard_hypotheses(ctx.handle, mine.ctypes.data_as(C.c_void_p), C.byref(count)))
    mine[count.value:] = np.nan
    payload = np.concatenate([mine.reshape(-1).view(np.int32).astype(np.int64), [count.value]]).astype(np.int64)
    every = nccl_allgather_wide(ctx, world, payload)
    hyps = []
    for r in range(world):
        cnt = int(every[r, -1])
        raw = every[r, :16].astype(np.int32).view(np.float32).reshape(4, 4)
        hyps.append([(raw[h, 0:2].tobytes(), raw[h, 2:4].tobytes()) for h in range(cnt)])
    picks, unresolved = resolve_psk_chain(hyps)
    state = np.zeros(2, dtype=np.float32)
    if rank < unresolved:
        ctx.check(lib.urh_costas_shard_adopt(ctx.handle, picks[rank], state.ctypes.data_as(C.c_void_p)))
    if unresolved < world:
        # hand-over from the first unresolved shard on: its predecessor's end state is known exactly
        carry = np.frombuffer(hyps[unresolved - 1][picks[unresolved - 1]][1], dtype=np.float32).copy()
        for turn in range(unresolved, world):
            out = np.zeros(2, dtype=np.float32)
            if turn == rank:
                ctx.check(lib.urh_costas_shard_resolve(ctx.handle, carry.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
                state = out
            allv = np.empty((world, 2), dtype=np.float32)
            ctx.check(lib.urh_nccl_allgather_host(ctx.handle, out.ctypes.data_as(C.c_void_p), allv.ctypes.data_as(C.c_void_p), out.nbytes))
            carry = allv[turn].copy()
    return state


def nccl_allgather_wide(ctx, world, values):
    """all-gather an int64 vector per rank over NCCL (pinned staging) -> array [world, len(values)]"""
    send = np.ascontiguousarray(values, dtype=np.int64)
    recv = np.empty((world, len(send)), dtype=np.int64)
    ctx.check(ctx.lib.urh_nccl_allgather_host(ctx.handle, send.ctypes.data_as(C.c_void_p), recv.ctypes.data_as(C.c_void_p), send.nbytes))
    return recv


def demod_digitize_sharded(ctx, hx, sb: ShardBuffer, global_offset, n_total, noise_mag, mod_type, center, tolerance,
                           samples_per_symbol, bits_per_symbol=1, center_spacing=0.1, d_qad=None, root=0):
    """FSK/ASK demod + digitize of a capture sharded over the ranks.  Returns the (k,2) pulse table on `root`
    (None elsewhere).  `d_qad` (optional DeviceArray[n_local]) receives this rank's demodulated samples."""
    lib = ctx.lib
    code = _lib.demod_mod_code(mod_type)
    summary = (C.c_int64 * 4)()
    ctx.check(lib.urh_shard_dense(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.dtype_code(sb.dtype), sb.n, int(hx.rank > 0),
                                  float(noise_mag), code, float(center), int(tolerance), int(bits_per_symbol), float(center_spacing),
                                  C.c_void_p(d_qad.ptr if d_qad is not None else 0), summary))
    mine = (int(summary[0]), int(summary[1]), int(summary[2]), int(summary[3]))
    every = hx.allgather(mine)
    carry = fold_carry([(c, l, w) for c, l, w, _ in every])[hx.rank]
    count = C.c_int64(0)
    d_pos, d_cls = C.c_void_p(), C.c_void_p()
    ctx.check(lib.urh_shard_candidates(ctx.handle, int(carry is not None), carry[0] if carry else 0, carry[1] if carry else 0,
                                       int(global_offset), C.byref(count), C.byref(d_pos), C.byref(d_cls), None))
    counts = hx.allgather(int(count.value))
    total = int(sum(counts))
    pos_all = cls_all = None
    if hx.rank == root:
        pos_all = DeviceArray(ctx, (max(total, 1),), np.int64)
        cls_all = DeviceArray(ctx, (max(total, 1),), np.int16)
    b8 = (C.c_int64 * hx.world)(*[c * 8 for c in counts])
    b2 = (C.c_int64 * hx.world)(*[c * 2 for c in counts])
    ctx.check(lib.urh_nccl_gatherv(ctx.handle, d_pos, C.c_void_p(pos_all.ptr if pos_all else 0), b8, root))
    ctx.check(lib.urh_nccl_gatherv(ctx.handle, d_cls, C.c_void_p(cls_all.ptr if cls_all else 0), b2, root))
    if hx.rank != root:
        ctx.sync()
        return None
    k = C.c_int64(0)
    ctx.check(lib.urh_pulses_from_table(ctx.handle, C.c_void_p(pos_all.ptr), C.c_void_p(cls_all.ptr), total, int(n_total),
                                        int(tolerance), code, int(samples_per_symbol), every[0][3], C.byref(k)))
    rows = np.empty((k.value, 2), dtype=np.int64)
    if k.value:
        ctx.check(lib.urh_fetch_pulses(ctx.handle, rows.ctypes.data_as(C.c_void_p), k.value))
    return rows


def nccl_allgather_i64(ctx, world, values):
    """all-gather a few int64 per rank over NCCL (device-staged, ~tens of microseconds) -> array [world, len(values)]"""
    send = np.ascontiguousarray(values, dtype=np.int64)
    recv = np.empty((world, len(send)), dtype=np.int64)
    entry = ctx.lib.urh_p2p_allgather_host if (getattr(ctx, "p2p", False) and send.nbytes <= 48) else ctx.lib.urh_nccl_allgather_host
    ctx.check(entry(ctx.handle, send.ctypes.data_as(C.c_void_p), recv.ctypes.data_as(C.c_void_p), send.nbytes))
    return recv


def previous_nonempty(values, counts, rank, default):
    """value of the nearest rank < `rank` whose count is non-zero, else `default`"""
    for r in range(rank - 1, -1, -1):
        if counts[r] > 0:
            return values[r]
    return default


def demod_digitize_distributed(ctx, rank, world, sb: ShardBuffer, global_offset, n_total, noise_mag, mod_type, center, tolerance,
                               samples_per_symbol, bits_per_symbol=1, center_spacing=0.1, d_qad=None, fetch=True, qad_source=None):
    """Sharded FSK/ASK demod + digitize with a DISTRIBUTED finish: no gather, every rank ends with the rows of its own
    shard (``merge_shard_rows`` joins them).  One library call per rank (urh_shard_digitize): the dense pass, then the
    tile-level finish whose three 16-byte exchanges (run carry / class of the last candidate / position of the last firing)
    are NCCL all-gathers enqueued on the context stream — the host waits once, for the row count.
    ``qad_source``: the shard is already demodulated (float32 DeviceArray) -> digitize from it instead of the IQ samples."""
    if os.environ.get("URH_B200_DIST_STEPWISE"):
        return demod_digitize_distributed_stepwise(ctx, rank, world, sb, global_offset, n_total, noise_mag, mod_type, center, tolerance,
                                                   samples_per_symbol, bits_per_symbol, center_spacing, d_qad, fetch, qad_source)
    lib = ctx.lib
    code = _lib.demod_mod_code(mod_type)
    k = C.c_int64(0)
    ctx.check(lib.urh_shard_digitize(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.dtype_code(sb.dtype),
                                     C.c_void_p(qad_source.ptr if qad_source is not None else 0), sb.n, int(rank > 0), float(noise_mag), code,
                                     float(center), int(tolerance), int(samples_per_symbol), int(bits_per_symbol), float(center_spacing),
                                     C.c_void_p(d_qad.ptr if d_qad is not None else 0), int(global_offset), int(n_total), C.byref(k)))
    if not fetch:
        return int(k.value)
    rows = np.empty((k.value, 2), dtype=np.int64)
    if k.value:
        ctx.check(lib.urh_fetch_pulses(ctx.handle, rows.ctypes.data_as(C.c_void_p), k.value))
    return rows


def demod_digitize_distributed_stepwise(ctx, rank, world, sb: ShardBuffer, global_offset, n_total, noise_mag, mod_type, center, tolerance,
                               samples_per_symbol, bits_per_symbol=1, center_spacing=0.1, d_qad=None, fetch=True, qad_source=None):
    """Call-by-call variant (host folds between the stages; kept as the reference for the one-call path).
    Sharded FSK/ASK demod + digitize with a DISTRIBUTED finish: no gather, every rank ends with the rows of its own
    shard (``merge_shard_rows`` joins them).  Three NCCL all-gathers of a few int64 per rank are the whole exchange:
      (last_cls, last_len, whole, init_cls)  ->  run carry into the shard;
      (candidate count, class of the last candidate)  ->  fire decision of the shard's first candidate;
      (firing count, position of the last firing)  ->  length of the shard's first pulse.
    ``qad_source``: the shard is already demodulated (float32 DeviceArray) -> digitize from it instead of the IQ samples."""
    lib = ctx.lib
    code = _lib.demod_mod_code(mod_type)
    summary = (C.c_int64 * 4)()
    if qad_source is not None:
        ctx.check(lib.urh_shard_dense_qad(ctx.handle, C.c_void_p(qad_source.ptr), sb.n, code, float(center), int(tolerance),
                                          int(bits_per_symbol), float(center_spacing), summary))
    else:
        ctx.check(lib.urh_shard_dense(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.dtype_code(sb.dtype), sb.n, int(rank > 0),
                                      float(noise_mag), code, float(center), int(tolerance), int(bits_per_symbol),
                                      float(center_spacing), C.c_void_p(d_qad.ptr if d_qad is not None else 0), summary))
    every = nccl_allgather_i64(ctx, world, list(summary))
    carry = fold_carry([(int(c), int(l), int(w)) for c, l, w, _ in every])[rank]
    init_cls = int(every[0][3])
    count, last_cls = C.c_int64(0), C.c_int(0)
    ctx.check(lib.urh_shard_candidates(ctx.handle, int(carry is not None), carry[0] if carry else 0, carry[1] if carry else 0,
                                       int(global_offset), C.byref(count), None, None, C.byref(last_cls)))
    cc = nccl_allgather_i64(ctx, world, [count.value, last_cls.value])
    prev_cls = int(previous_nonempty(cc[:, 1], cc[:, 0], rank, init_cls))
    fired, last_pos = C.c_int64(0), C.c_int64(-1)
    ctx.check(lib.urh_shard_fire(ctx.handle, prev_cls, C.byref(fired), C.byref(last_pos)))
    ff = nccl_allgather_i64(ctx, world, [fired.value, last_pos.value])
    prev_fired = int(previous_nonempty(ff[:, 1], ff[:, 0], rank, -1))
    k = C.c_int64(0)
    ctx.check(lib.urh_shard_rows(ctx.handle, int(n_total), int(tolerance), code, int(samples_per_symbol), prev_fired,
                                 int(rank == world - 1), C.byref(k)))
    if not fetch:
        return int(k.value)
    rows = np.empty((k.value, 2), dtype=np.int64)
    if k.value:
        ctx.check(lib.urh_fetch_pulses(ctx.handle, rows.ctypes.data_as(C.c_void_p), k.value))
    return rows


def center_protocol(rank, world, kept, window_stats, histogram, allgather_i64, allreduce_sum_i64, max_size=None):
    """The exchange behind a capture-wide detect_center, independent of where the numbers come from (GPU + NCCL in
    ``detect_center_distributed``; numpy + gloo in tests/test_dist_cpu.py).
      kept                      number of samples this shard keeps (qad > -4)
      window_stats(lr0, lr1)    -> [count, min, max, sum, sumsq] of the kept samples of LOCAL rank [lr0, lr1)
      histogram(lr0, lr1, edges)-> int64 counts of those samples for np.histogram(…, bins=edges)
      allgather_i64(values)     -> array [world, len(values)]; allreduce_sum_i64(array) -> array
    Every rank returns the same center (or None)."""
    from .ainterpretation.AutoInterpretation import center_rank_window, center_stats_from_window, pick_center_from_histogram, \
        center_bin_edges
    counts = np.asarray(allgather_i64([int(kept)]))[:, 0]
    total, offset = int(counts.sum()), int(counts[:rank].sum())
    r0, r1 = center_rank_window(total, max_size)
    lr0 = min(max(r0 - offset, 0), int(kept))
    lr1 = min(max(r1 - offset, 0), int(kept))
    w = np.ascontiguousarray(window_stats(lr0, lr1), dtype=np.float64)
    parts = np.asarray(allgather_i64(w.view(np.int64))).view(np.float64)
    g = np.array([parts[:, 0].sum(), parts[:, 1].min(), parts[:, 2].max(), 0.0, 0.0])
    for q in range(world):  # rank order, so every rank (and every world size's replay) folds identically
        g[3] += parts[q, 3]
        g[4] += parts[q, 4]
    st = center_stats_from_window(total, r0, r1, g)
    edges = center_bin_edges(st)
    if edges is None:
        return None
    y = allreduce_sum_i64(np.ascontiguousarray(histogram(lr0, lr1, edges), dtype=np.int64))
    return pick_center_from_histogram(y, edges)


def detect_center_distributed(ctx, rank, world, sb: ShardBuffer, noise_mag, mod_type, d_qad, max_size=None):
    """afp_demod of the shard into ``d_qad`` + the capture-wide detect_center, every rank ending with the same center.
    Exchange (NCCL, a few hundred bytes + one histogram): kept-sample counts -> global rank window; per-rank window
    partials {count, min, max, sum, sumsq} folded in rank order (deterministic) -> bin edges; histogram all-reduce."""
    from .device import DeviceArray
    lib = ctx.lib
    code = _lib.demod_mod_code(mod_type)
    kept = C.c_int64(0)
    ctx.check(lib.urh_afp_demod_tiles(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.dtype_code(sb.dtype), sb.n, float(noise_mag), code,
                                      C.c_void_p(d_qad.ptr), int(rank > 0), C.byref(kept)))

    def window_stats(lr0, lr1):
        w = np.zeros(5, dtype=np.float64)
        ctx.check(lib.urh_center_window_stats(ctx.handle, C.c_void_p(d_qad.ptr), sb.n, lr0, lr1, w.ctypes.data_as(C.c_void_p)))
        return w

    def histogram(lr0, lr1, edges):
        nbins = len(edges) - 1
        y = np.zeros(nbins, dtype=np.int64)
        ctx.check(lib.urh_center_histogram_tiles(ctx.handle, C.c_void_p(d_qad.ptr), sb.n, lr0, lr1, C.c_double(edges[0]),
                                                 C.c_double(edges[1] - edges[0]), nbins, y.ctypes.data_as(C.c_void_p)))
        return y

    def allreduce(y):
        y = np.ascontiguousarray(y, dtype=np.int64)
        ctx.check(lib.urh_nccl_allreduce_host_i64(ctx.handle, y.ctypes.data_as(C.c_void_p), len(y), 0))
        return y

    return center_protocol(rank, world, kept.value, window_stats, histogram, lambda v: nccl_allgather_i64(ctx, world, v), allreduce,
                           max_size)


def demod_center_digitize_distributed(ctx, rank, world, sb: ShardBuffer, global_offset, n_total, noise_mag, mod_type, tolerance,
                                      samples_per_symbol, d_qad, bits_per_symbol=1, center_spacing=0.1, max_size=None, fetch=True,
                                      host_iq=None, rows_out=None, chunk_samples=1 << 24):
    """BASELINE configs[1]/[4] on N GPUs: demod + capture-wide detect_center + digitize of ONE sharded capture, one library
    call per rank (urh_shard_demod_center_digitize).  Exchanges, all NCCL on the context stream with device buffers: kept
    counts (8 B), window partials (32 B), the histogram all-reduce, then the digitizer's three 16-byte all-gathers.
    The digitizer pass reads the shard's qad (4 B/sample) once the center is known.  -> (center, rows or count).
    ``host_iq``: this rank's shard in (pinned) host memory: it is streamed into ``sb.shard`` in chunks while the chunks that have
    landed are demodulated (the halo sample must already be in ``sb.halo``); ``rows_out``: pinned int64 buffer for the rows."""
    lib = ctx.lib
    code = _lib.demod_mod_code(mod_type)
    if bits_per_symbol == 1 and not os.environ.get("URH_B200_DIST_STEPWISE"):
        center, state, k = C.c_double(0.0), C.c_int(0), C.c_int64(0)
        if host_iq is not None:
            ctx.check(lib.urh_shard_demod_center_digitize_host(ctx.handle, host_iq.ctypes.data_as(C.c_void_p), _lib.dtype_code(sb.dtype), sb.n,
                                                               int(rank > 0), float(noise_mag), code, int(tolerance), int(samples_per_symbol),
                                                               -1 if max_size is None else int(max_size), int(chunk_samples),
                                                               C.c_void_p(sb.shard.ptr), C.c_void_p(d_qad.ptr), int(global_offset), int(n_total),
                                                               C.byref(center), C.byref(state), C.byref(k)))
        else:
            ctx.check(lib.urh_shard_demod_center_digitize(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.dtype_code(sb.dtype), sb.n, int(rank > 0),
                                                          float(noise_mag), code, int(tolerance), int(samples_per_symbol),
                                                          -1 if max_size is None else int(max_size), C.c_void_p(d_qad.ptr), int(global_offset),
                                                          int(n_total), C.byref(center), C.byref(state), C.byref(k)))
        if state.value == 0:
            return None, (np.zeros((0, 2), dtype=np.int64) if fetch else 0)
        if state.value == 1:
            if not fetch:
                return float(center.value), int(k.value)
            if rows_out is not None and rows_out.dtype == np.int64 and rows_out.size >= 2 * k.value:
                rows = rows_out.reshape(-1)[: 2 * k.value].reshape(k.value, 2)
            else:
                rows = np.empty((k.value, 2), dtype=np.int64)
            if k.value:
                ctx.check(lib.urh_fetch_pulses(ctx.handle, rows.ctypes.data_as(C.c_void_p), k.value))
            return float(center.value), rows
        # state 2: a tie the device must not break (every rank sees the same histogram, so every rank lands here together)
    center = detect_center_distributed(ctx, rank, world, sb, noise_mag, mod_type, d_qad, max_size)
    if center is None:
        return None, (np.zeros((0, 2), dtype=np.int64) if fetch else 0)
    out = demod_digitize_distributed(ctx, rank, world, sb, global_offset, n_total, noise_mag, mod_type, float(center), tolerance,
                                     samples_per_symbol, bits_per_symbol, center_spacing, None, fetch, qad_source=d_qad)
    return center, out


def merge_shard_rows(parts):
    """concatenate per-shard pulse tables; equal states that meet at a shard edge are one pulse (pyx:475-476)"""
    out = []
    for rows in parts:
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 2)
        if len(rows) == 0:
            continue
        if out and out[-1][-1, 0] == rows[0, 0]:
            out[-1][-1, 1] += rows[0, 1]
            rows = rows[1:]
        if len(rows):
            out.append(rows.copy())
    return np.concatenate(out) if out else np.zeros((0, 2), dtype=np.int64)


def detect_noise_level_sharded(ctx, hx, sb: ShardBuffer, global_offset, n_total):
    """AutoInterpretation.detect_noise_level over the whole capture: per-rank partial (sum, max) of the 100 global,
    end-aligned chunks, ONE NCCL all-reduce (sum half / max half share a buffer: max via sum of ... two calls), then
    the reference's host logic."""
    from .ainterpretation import AutoInterpretation as AI

    if n_total <= 3:
        return 0
    chunksize, nchunks = AI._chunking(n_total)
    # global chunk j covers [n_total-(j+1)*cs, n_total-j*cs); intersect with this shard
    sums = np.zeros(nchunks, dtype=np.float64)
    maxs = np.full(nchunks, -1.0, dtype=np.float64)
    lo, hi = int(global_offset), int(global_offset) + sb.n
    for j in range(nchunks):
        c0, c1 = n_total - (j + 1) * chunksize, n_total - j * chunksize
        a, b = max(c0, lo), min(c1, hi)
        if a >= b:
            if c1 <= lo:
                break
            continue
        part = sb.shard[a - lo: b - lo]
        s1, m1 = np.zeros(1), np.zeros(1)
        ctx.check(ctx.lib.urh_noise_chunk_stats_iq(ctx.handle, C.c_void_p(part.ptr), _lib.dtype_code(sb.dtype), b - a, b - a, 1,
                                                   s1.ctypes.data_as(C.c_void_p), m1.ctypes.data_as(C.c_void_p)))
        sums[j], maxs[j] = s1[0], m1[0]
    d_s = DeviceArray(ctx, (nchunks,), np.float64).set(sums)
    d_m = DeviceArray(ctx, (nchunks,), np.float64).set(maxs)
    ctx.check(ctx.lib.urh_nccl_allreduce_f64(ctx.handle, C.c_void_p(d_s.ptr), nchunks, 0))
    ctx.check(ctx.lib.urh_nccl_allreduce_f64(ctx.handle, C.c_void_p(d_m.ptr), nchunks, 1))
    return AI._noise_from_chunk_stats(n_total, chunksize, d_s.get(), d_m.get(), np.float64)


# ---- AutoInterpretation.estimate over a sharded capture (BASELINE configs[4]) ------------------------------------------------------
def segment_messages_sharded(ctx, hx, d_mag, global_offset, n_total, noise_threshold):
    """segment_messages_from_magnitudes (auto_interpretation.pyx:55-111) of a capture whose magnitudes are spread over the ranks:
    every rank runs the dense pass over its shard; the closing-run summaries are folded (``fold_carry``) so that runs crossing a
    shard edge count their samples on both sides; the run tables (a few entries per message) are concatenated in rank order and
    every rank runs the reference's two-state machine on them.  Returns the same list of (start, end) on every rank."""
    lib = ctx.lib
    n_local = len(d_mag)
    summary = (C.c_int64 * 4)()
    ctx.check(lib.urh_segment_shard_pass(ctx.handle, C.c_void_p(d_mag.ptr), int(d_mag.dtype == np.float64), n_local,
                                         float(noise_threshold), summary))
    every = hx.allgather((int(summary[0]), int(summary[1]), int(summary[2]), int(summary[3])))
    carries = fold_carry([(c, l, w) for c, l, w, _ in every])
    carry = carries[hx.rank]
    count = C.c_int64(0)
    ctx.check(lib.urh_shard_candidates(ctx.handle, int(carry is not None), carry[0] if carry else 0, carry[1] if carry else 0,
                                       int(global_offset), C.byref(count), None, None, None))
    pos = np.empty(count.value, dtype=np.int64)
    cls = np.empty(count.value, dtype=np.int16)
    if count.value:
        ctx.check(lib.urh_fetch_candidates(ctx.handle, pos.ctypes.data_as(C.c_void_p), cls.ctypes.data_as(C.c_void_p), count.value))
    tables = hx.allgather((pos, cls))
    pos_all = np.ascontiguousarray(np.concatenate([t[0] for t in tables]))
    cls_all = np.ascontiguousarray(np.concatenate([t[1] for t in tables]))
    # the run that ends the capture: fold every shard's closing run
    last_cls, last_len = None, 0
    for c, l, w, _ in every:
        if last_cls is not None and w and c == last_cls:
            last_len += l
        else:
            last_cls, last_len = c, l
    cap = max(16, len(pos_all) + 2)
    seg = np.empty((cap, 2), dtype=np.int64)
    k = C.c_int64(0)
    ctx.check(lib.urh_segments_from_runs(pos_all.ctypes.data_as(C.c_void_p), cls_all.ctypes.data_as(C.c_void_p), len(pos_all), int(every[0][3]),
                                         int(last_cls), int(last_len), int(n_total), seg.ctypes.data_as(C.c_void_p), cap, C.byref(k)))
    return [(int(a), int(b)) for a, b in seg[: k.value]]


def fetch_range(ctx, rank, bounds, d_local, g0, g1, owner):
    """Collective: every rank calls it with the same (g0, g1, owner).  `d_local` is this rank's shard (1-D or (n, 2) DeviceArray) of
    a capture cut at `bounds` [(start, end) per rank].  Returns on `owner` a DeviceArray with elements [g0, g1) of the capture
    (its own part copied, the other parts received over NCCL); None elsewhere."""
    lib = ctx.lib
    item = d_local.nbytes // max(1, len(d_local))
    out = None
    if rank == owner:
        shape = (g1 - g0,) + tuple(d_local.shape[1:])
        out = DeviceArray(ctx, shape, d_local.dtype)
    for q, (a, b) in enumerate(bounds):
        lo, hi = max(a, g0), min(b, g1)
        if lo >= hi:
            continue
        nbytes = (hi - lo) * item
        if q == owner:
            if rank == owner:
                ctx.check(lib.urh_memcpy_d2d(ctx.handle, C.c_void_p(out.ptr + (lo - g0) * item), C.c_void_p(d_local.ptr + (lo - a) * item), nbytes))
        elif rank == q:
            ctx.check(lib.urh_nccl_sendrecv(ctx.handle, C.c_void_p(d_local.ptr + (lo - a) * item), nbytes, owner, None, 0, -1))
        elif rank == owner:
            ctx.check(lib.urh_nccl_sendrecv(ctx.handle, None, 0, -1, C.c_void_p(out.ptr + (lo - g0) * item), nbytes, q))
    return out


def estimate_sharded(ctx, hx, sb: ShardBuffer, bounds, n_total, noise=None, modulation=None):
    """AutoInterpretation.estimate (AutoInterpretation.py:373-471) of ONE capture sharded by contiguous sample range: the same dict
    on every rank, equal to the single-GPU / reference result.
      noise       global end-aligned chunk statistics, NCCL all-reduce (detect_noise_level_sharded)
      messages    sharded segmentation with run carry (segment_messages_sharded)
      modulation  detect_modulation on the first 100 messages, each on the rank that owns its start
      demod       ASK / FSK with the 1-sample halo; PSK with the speculative Costas loop over shards (afp_demod_psk_sharded)
      parameters  detect_center / plateau lengths / tolerance / bit length per message on the owning rank (a message that
                  straddles a shard edge is completed over NCCL), gathered and reduced exactly as the reference does."""
    from .ainterpretation import AutoInterpretation as AI
    from .cythonext import auto_interpretation as c_ai
    from .cythonext import signal_functions as sf
    from .cythonext import util

    rank, world = hx.rank, hx.world
    lo, hi = bounds[rank]
    d_mag = util.get_magnitudes(sb.shard)
    if noise is None:
        noise = detect_noise_level_sharded(ctx, hx, sb, lo, n_total)
    message_indices = segment_messages_sharded(ctx, hx, d_mag, lo, n_total, noise)
    d_mag.free()

    def owner_of(start):
        for q, (a, b) in enumerate(bounds):
            if a <= start < b:
                return q
        return world - 1

    def message_slice(d_shard, start, end):
        """[start, end) of the capture on the rank owning `start` (collective when the message leaves that rank's shard)"""
        own = owner_of(start)
        a, b = bounds[own]
        if end <= b:
            return (d_shard[start - a: end - a] if rank == own else None), own
        return fetch_range(ctx, rank, bounds, d_shard, start, end, own), own

    if modulation is None:
        found_local = []
        for idx, (start, end) in enumerate(message_indices[0:100]):
            part, own = message_slice(sb.shard, start, end)
            if rank == own:
                from .signalprocessing.IQArray import IQArray
                mod = AI.detect_modulation(IQArray(np.ascontiguousarray(part.get()), _owned=True).as_complex64())
                if mod is not None:
                    found_local.append((idx, mod))
        found = sorted(x for part in hx.allgather(found_local) for x in part)
        modulation = AI.most_common([m for _, m in found]) if found else None
    if modulation is None:
        return None
    if modulation == "OOK":
        message_indices = AI.merge_message_segments_for_ook(message_indices)
    d_qad = DeviceArray(ctx, (hi - lo,), np.float32)
    if modulation == "PSK":
        afp_demod_psk_sharded(ctx, rank, world, sb, noise, 2, 0.1, d_qad)
    else:
        mt = "ASK" if modulation in ("OOK", "ASK") else "FSK"
        kept = C.c_int64(0)
        ctx.check(ctx.lib.urh_afp_demod_tiles(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.dtype_code(sb.dtype), sb.n, float(noise),
                                              _lib.demod_mod_code(mt), C.c_void_p(d_qad.ptr), int(rank > 0), C.byref(kept)))
    local = []   # (message index, center, bit_length or None, tolerance or None)
    for idx, (start, end) in enumerate(message_indices):
        msg, own = message_slice(d_qad, int(start), int(end))
        if rank != own:
            continue
        center = AI.detect_center(msg)
        if center is None:
            continue
        plateau_lengths = c_ai.get_plateau_lengths(msg, center, percentage=25)
        tolerance = AI.estimate_tolerance_from_plateau_lengths(plateau_lengths)
        tol_entry = None
        if tolerance is None:
            tolerance = 0
        else:
            tol_entry = tolerance
        merged = AI.merge_plateau_lengths(plateau_lengths, tolerance=tolerance)
        bit_entry = None
        if len(merged) >= 2:
            bit_length = AI.get_bit_length_from_plateau_lengths(merged)
            if bit_length > tolerance + 1:
                bit_entry = (float(center), bit_length)
        local.append((idx, tol_entry, bit_entry))
    rows = sorted(x for part in hx.allgather(local) for x in part)
    tolerances = [t for _, t, _ in rows if t is not None]
    centers = [b[0] for _, _, b in rows if b is not None]
    bit_lengths = [b[1] for _, _, b in rows if b is not None]
    if modulation in ("OOK", "ASK"):
        center = AI.min_without_outliers(np.array(centers), z=2)
        if center is None:
            return None
    elif len(centers) > 0:
        center = np.mean(centers)
    else:
        return None
    bit_length = AI.get_most_frequent_value(bit_lengths)
    if bit_length is None:
        return None
    try:
        tolerance = np.percentile(tolerances, 50)
    except IndexError:
        tolerance = max(1, int(0.05 * bit_length))
    return {"modulation_type": "ASK" if modulation == "OOK" else modulation, "bit_length": bit_length, "center": center,
            "tolerance": int(tolerance), "noise": noise}
