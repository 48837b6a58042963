#!/usr/bin/env python
"""BASELINE configs[4]: a BPSK capture of N x 2^log2n samples sharded across N GPUs (contiguous ranges, Costas halo), demodulated by
the speculative Costas loop over shards, then AutoInterpretation.estimate over the sharded capture (modulation given, then detected).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_config4.py --log2n 30
    (N = 1: python tools/bench_config4.py)

Parity at size, on every run: rank r checks the first 2^22 demodulated samples of its shard against the CPU oracle's Costas loop
started from the loop state the preceding shard ended in (oracle.costas_from) — bit for bit; rank 0 starts from the reference's
initial state.  Prints one JSON line (rank 0)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=30)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--noise", type=float, default=0.3)
    ap.add_argument("--no-estimate", action="store_true")
    ap.add_argument("--given-only", action="store_true",
                    help="estimate with modulation='PSK' only (with modulation=None the reference's detector calls this capture FSK and its "
                         "Python plateau loops then walk millions of one-sample plateaus per message: minutes of host time at 2^33 samples)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from urh_b200 import _lib, dist as udist
    from urh_b200.device import DeviceArray

    ctx = _lib.default_context(local_rank)
    lib = ctx.lib
    n = 1 << args.log2n
    n_total = n * world
    bounds = [(q * n, (q + 1) * n) for q in range(world)]
    halo = int(lib.urh_costas_halo_samples())
    sb = udist.ShardBuffer(ctx, n, np.float32, halo=halo)
    sps, order = 300, 2
    period, burst = 3_000_000, 2_400_000
    ctx.check(lib.urh_synth_psk(ctx.handle, C.c_void_p(sb.shard.ptr), n, rank * n, sps, order, C.c_double(0.025), 1.0, 0.05, 4242, period, burst,
                                int(0.97 * n_total)))
    ctx.sync()

    class Solo(object):
        rank, world = 0, 1

        def allgather(self, obj):
            return [obj]

        def barrier(self):
            pass

    hx = Solo()
    if world > 1:
        hx = udist.HostExchange()
        udist.init_nccl(ctx, hx)
        udist.exchange_halo(ctx, hx, sb)
    d_qad = DeviceArray(ctx, (n,), np.float32)

    def demod():
        if world > 1:
            return udist.afp_demod_psk_sharded(ctx, rank, world, sb, args.noise, order, 0.1, d_qad)
        ctx.check(lib.urh_afp_demod(ctx.handle, C.c_void_p(sb.shard.ptr), _lib.DT_F32, n, args.noise, _lib.MOD_PSK, order, 0.1, C.c_void_p(d_qad.ptr)))
        return None

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()

    for _ in range(2):
        state = demod()
    barrier()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        state = demod()
    ms = ctx.timer_stop()
    ms = max(ms, (time.perf_counter() - t0) * 1e3) / args.steps
    if world > 1:
        import torch

        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    redone = int(lib.urh_costas_last_redone(ctx.handle))

    # ---- parity at size: the head of every shard against the oracle's loop continued from the preceding shard's end state --------
    from oracle import oracle

    oracle.build()
    w = min(n, 1 << 22)
    states = hx.allgather(None if state is None else [float(state[0]), float(state[1])])
    head = sb.shard[0:w].get()
    got = d_qad[0:w].get()
    if rank == 0:
        ref = oracle.afp_demod(head, args.noise, "PSK", order)
        bad = int(np.count_nonzero(got[1:].view(np.uint32) != ref[1:].view(np.uint32)))
    else:
        ref, _ = oracle.costas_from(head, args.noise, order, states[rank - 1])
        bad = int(np.count_nonzero(got.view(np.uint32) != ref.view(np.uint32)))
    bads = hx.allgather(bad)

    est = {}
    if not args.no_estimate:
        for given in (("PSK",) if args.given_only else ("PSK", None)):
            barrier()
            t0 = time.perf_counter()
            e = udist.estimate_sharded(ctx, hx, sb, bounds, n_total, noise=None, modulation=given) if world > 1 else None
            if world == 1:
                from urh_b200.ainterpretation import AutoInterpretation as AI
                from urh_b200.signalprocessing.IQArray import IQArray
                e = AI.estimate(IQArray(sb.shard.get(), _owned=True), noise=None, modulation=given)
            barrier()
            est["modulation=%s" % given] = {"seconds": time.perf_counter() - t0,
                                            "result": None if e is None else {k: (float(v) if not isinstance(v, str) else v) for k, v in e.items()}}
    if rank == 0:
        print(json.dumps({"config": "configs[4]: BPSK capture of %d x 2^%d samples sharded over %d GPU(s), Costas halo %d samples"
                                    % (world, args.log2n, world, halo),
                          "n_gpus": world, "samples": n_total, "psk_demod_ms": ms, "MSamples_per_s": n_total / ms / 1e3,
                          "noise_mag": args.noise, "super_chunks_redone_rank0": redone,
                          "parity": {"window_samples_per_shard": w, "words_differing_per_rank": bads, "ok": all(b == 0 for b in bads)},
                          "estimate": est}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
