// grab_pulse_lens after the dense pass, entirely at TILE level (signal_functions.pyx:455-495).
//
// The dense pass leaves per tile a 16-byte summary and the staged interior candidates.  Everything that used to
// work on the gathered candidate table (gather, fire flags, three int64 scans over ~n/100 entries, three scalar
// read-backs) is restated on the tile table (n/2048 entries) as three look-back scans and one row kernel:
//
//   A  run carry           exclusive scan of {class, length, whole?} of each tile's closing run  -> carry[t]
//   B  candidates          per tile: head candidate from carry[t] (+ the carry of the preceding shards), number of
//                          candidates, class of the last one; scan -> class of the candidate preceding the tile
//   C  firings             per tile: walk its candidates (a candidate fires iff its class differs from the one
//                          before it), count firings, position of the last; scan -> row offset, previous firing
//   D  rows                per tile: walk again, write (state, length) rows at the tile's row offset; the last
//                          tile appends the tail row (pyx:485-493) and the row count
// (C and D run one THREAD per tile: a tile holds ~20 candidates, and 32 independent walks per warp keep far more loads in
// flight than one warp per tile did — measured 42 + 104 us against 113 + 128 us at 2^19 tiles.)
//
// One read-back (row count) ends the call.  Rows go straight into the context's pulse buffer, sized optimistically;
// an overflow only repeats stage D.  ASK (short pauses relabelled, pyx:471-473, so equal neighbours can meet) runs a
// fourth scan that merges equal neighbours.
//
// Sharded captures: between the stages every rank publishes its scan total (16 bytes) with an NCCL
// all-gather ON THE CONTEXT STREAM into device memory and a one-thread kernel folds the totals of the preceding
// ranks; the host never waits between the stages.
#include "sparse.cuh"
#include "tilescan.cuh"

#include <limits.h>

#define CLS_NONE INT_MIN

// ---- scan elements --------------------------------------------------------------------------------------------------
struct __align__(16) CandAgg {
    int64_t cnt;       // candidates
    int32_t last_cls;  // class of the last candidate (CLS_NONE: no candidate in the span)
    int32_t pad;
};
struct CandOp {
    __device__ __forceinline__ CandAgg operator()(const CandAgg& a, const CandAgg& b) const {
        CandAgg r;
        r.cnt = a.cnt + b.cnt;
        r.last_cls = (b.last_cls != CLS_NONE) ? b.last_cls : a.last_cls;
        r.pad = 0;
        return r;
    }
};
struct __align__(16) FireAgg {
    int64_t fired;     // firings
    int64_t last_pos;  // global position of the last firing (-1: none in the span)
};
struct FireOp {
    __device__ __forceinline__ FireAgg operator()(const FireAgg& a, const FireAgg& b) const {
        FireAgg r;
        r.fired = a.fired + b.fired;
        r.last_pos = (b.last_pos >= 0) ? b.last_pos : a.last_pos;
        return r;
    }
};

// ---- A ---------------------------------------------------------------------------------------------------------------
struct ScanRunCarry {
    const UrhTileSummary* tiles;
    int64_t n;
    RunCarry* carry;
    __device__ __forceinline__ RunCarry load(int64_t t) const {
        const int64_t rem = n - t * URH_TILE;
        const int tile_len = rem < URH_TILE ? (int)rem : URH_TILE;
        const UrhTileSummary s = tiles[t];
        RunCarry r;
        r.len = s.tail_len;
        r.cls = s.last_cls;
        r.flags = (s.head_len == tile_len) ? 1 : 0;
        return r;
    }
    __device__ __forceinline__ void post(int64_t t, const RunCarry& excl, const RunCarry&) const { carry[t] = excl; }
};

// ---- B ---------------------------------------------------------------------------------------------------------------
struct ScanCandidates {
    const UrhTileSummary* tiles;
    const uint32_t* staging;
    int stage_cap;
    const RunCarry* carry;
    const RunCarry* xcarry;   // device: the run that ends right before this shard (nullptr: unsharded)
    int tol;
    int32_t* head_rel;
    int32_t* prev_cls;
    __device__ __forceinline__ CandAgg load(int64_t t) const {
        const UrhTileSummary s = tiles[t];
        RunCarry c = carry[t];
        if (xcarry) c = RunCarryOp()(*xcarry, c);
        int64_t start_len = 0;
        if (!(c.flags & 2) && c.cls == s.first_cls) start_len = c.len;
        int32_t rel = -1;
        if (start_len <= tol && (int64_t)tol < start_len + s.head_len) rel = (int32_t)(tol - start_len);
        head_rel[t] = rel;
        CandAgg r;
        r.cnt = (int64_t)s.ncand + (rel >= 0 ? 1 : 0);
        r.last_cls = CLS_NONE;
        if (s.ncand > 0) r.last_cls = (int32_t)(staging[t * (int64_t)stage_cap + s.ncand - 1] & 0xffffu) - 1;
        else if (rel >= 0) r.last_cls = s.first_cls;
        r.pad = 0;
        return r;
    }
    __device__ __forceinline__ void post(int64_t t, const CandAgg& excl, const CandAgg&) const { prev_cls[t] = excl.last_cls; }
};

// ---- C ---------------------------------------------------------------------------------------------------------------
struct ScanFirings {
    const UrhTileSummary* tiles;
    const uint32_t* staging;
    int stage_cap;
    const int32_t* head_rel;
    const int32_t* prev_cls;
    const int16_t* d_prev0;   // device: class of the candidate preceding the shard (the digitizer's initial state)
    int64_t global_offset;
    int64_t* row_off;
    int64_t* prev_fired;
    __device__ __forceinline__ FireAgg load(int64_t t) const {
        const UrhTileSummary s = tiles[t];
        int prev = prev_cls[t];
        if (prev == CLS_NONE) prev = *d_prev0;
        const int64_t base = t * URH_TILE + global_offset;
        FireAgg r;
        r.fired = 0;
        r.last_pos = -1;
        const int32_t rel = head_rel[t];
        if (rel >= 0) {
            const int c = s.first_cls;
            if (c != prev) { r.fired++; r.last_pos = base + rel; }
            prev = c;
        }
        const uint32_t* st = staging + t * (int64_t)stage_cap;
        for (int j = 0; j < s.ncand; j++) {
            const uint32_t v = st[j];
            const int c = (int)(v & 0xffffu) - 1;
            if (c != prev) { r.fired++; r.last_pos = base + (v >> 16); }
            prev = c;
        }
        return r;
    }
    __device__ __forceinline__ void post(int64_t t, const FireAgg& excl, const FireAgg&) const {
        row_off[t] = excl.fired;
        prev_fired[t] = excl.last_pos;
    }
};

// ---- D ---------------------------------------------------------------------------------------------------------------
// out = (state, length) pairs; rows beyond cap_rows are dropped (the caller grows the buffer and repeats).
// d_out[0] = rows written incl. tail, d_out[1] = firings.
__global__ void __launch_bounds__(256) k_finish_rows(const UrhTileSummary* __restrict__ tiles, const uint32_t* __restrict__ staging,
                                                    int stage_cap, const int32_t* __restrict__ head_rel, const int32_t* __restrict__ prev_cls,
                                                    const int16_t* __restrict__ d_prev0, const int64_t* __restrict__ row_off,
                                                    const int64_t* __restrict__ prev_fired, const int64_t* __restrict__ d_xprev_fired,
                                                    int64_t ntiles, int64_t global_offset, int64_t n_total, int tol, int is_ask, int64_t sps,
                                                    int emit_tail, int64_t* __restrict__ out, int64_t cap_rows, int64_t* __restrict__ d_out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const UrhTileSummary s = tiles[t];
    int prev = prev_cls[t];
    if (prev == CLS_NONE) prev = *d_prev0;
    int64_t pp = prev_fired[t];
    if (pp < 0) pp = *d_xprev_fired;
    int64_t idx = row_off[t];
    const int64_t base = t * URH_TILE + global_offset;
    auto fire = [&](int64_t p, int c) {
        if (c != prev) {
            // pulse lengths (pyx:476-482): the first pulse of the capture is counted from its start
            const int64_t rec = (pp >= 0) ? (p - pp) : (p + 1 - tol);
            int64_t st = prev;
            if (is_ask && st == -1 && rec < sps) st = 0;   // ASK: a pause shorter than one symbol is a zero (pyx:471-473)
            if (idx < cap_rows) *((longlong2*)out + idx) = make_longlong2(st, rec);   // one 16-byte store per row
            idx++;
            pp = p;
        }
        prev = c;
    };
    const int32_t rel = head_rel[t];
    if (rel >= 0) fire(base + rel, s.first_cls);
    const uint32_t* st = staging + t * (int64_t)stage_cap;
    for (int j = 0; j < s.ncand; j++) {
        const uint32_t v = st[j];
        fire(base + (v >> 16), (int)(v & 0xffffu) - 1);
    }
    if (t == ntiles - 1) {
        const int64_t fired = idx;
        // tail row (pyx:485-493): appended only while fewer than n rows exist
        if (emit_tail && (is_ask || fired < n_total)) {
            if (idx < cap_rows) *((longlong2*)out + idx) = make_longlong2(prev, (pp >= 0) ? (n_total - 1 - pp) : (n_total - tol));
            idx++;
        }
        d_out[0] = idx;
        d_out[1] = fired;
    }
}

// ---- ASK: merge equal neighbours (pyx:475-476) -------------------------------------------------------------------------
struct ScanMergeRows {
    const int64_t* raw;    // (state, length) x rows, the tail row last when has_tail
    int64_t rows;
    int64_t n_total;
    int has_tail;
    int64_t* out;
    int64_t* d_k;
    __device__ __forceinline__ int64_t load(int64_t r) const { return (r == 0 || raw[2 * r] != raw[2 * r - 2]) ? 1 : 0; }
    __device__ __forceinline__ void post(int64_t r, const int64_t& excl, const int64_t& head) const {
        const bool is_tail = has_tail && r == rows - 1;
        // the tail row is appended only while fewer than n (merged) rows exist (pyx:487)
        if (is_tail && excl >= n_total) {
            *d_k = excl;
            return;
        }
        const int64_t o = excl + head - 1;
        if (head) out[2 * o] = raw[2 * r];
        atomicAdd((unsigned long long*)&out[2 * o + 1], (unsigned long long)raw[2 * r + 1]);
        if (r == rows - 1) *d_k = o + 1;
    }
};
struct AddI64 {
    __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a + b; }
};

// ---- folding the totals of the preceding ranks (sharded captures) ------------------------------------------------------------
// Stage-1 message of a rank: {RunCarry total (2 x int64), init class, pad}; stages 2 and 3: the CandAgg / FireAgg total.
__global__ void k_pack_stage1(const int16_t* __restrict__ d_init, const RunCarry* __restrict__ total, int64_t* __restrict__ msg) {
    memcpy(msg, total, sizeof(RunCarry));
    msg[2] = *d_init;
    msg[3] = 0;
}
__global__ void k_fold_carry(const int64_t* __restrict__ all, int rank, RunCarry* __restrict__ xcarry) {
    RunCarry acc;
    acc.len = 0; acc.cls = 0; acc.flags = 2 | 1;
    for (int q = 0; q < rank; q++) {
        RunCarry c;
        memcpy(&c, all + 4 * q, sizeof(c));
        acc = RunCarryOp()(acc, c);
    }
    if (rank == 0) acc.flags = 2;   // nothing precedes the first shard
    else acc.flags &= ~1;           // the incoming run is never "the whole span" of this shard
    *xcarry = acc;
}
__global__ void k_fold_prev_cls(const int64_t* __restrict__ all, int rank, const int64_t* __restrict__ stage1, int16_t* __restrict__ prev0) {
    int v = (int)stage1[2];   // rank 0's initial class
    for (int q = 0; q < rank; q++) {
        CandAgg c;
        memcpy(&c, all + 2 * q, sizeof(c));
        if (c.last_cls != CLS_NONE) v = c.last_cls;
    }
    *prev0 = (int16_t)v;
}
__global__ void k_fold_prev_fired(const int64_t* __restrict__ all, int rank, int64_t* __restrict__ xprev) {
    int64_t v = -1;
    for (int q = 0; q < rank; q++) {
        FireAgg c;
        memcpy(&c, all + 2 * q, sizeof(c));
        if (c.last_pos >= 0) v = c.last_pos;
    }
    *xprev = v;
}

int urh_coll_allgather(urh_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank);   // nccl.cu: mailboxes or NCCL
extern "C" int urh_p2p_check(urh_ctx* ctx);

// ---- driver --------------------------------------------------------------------------------------------------------------
struct FinishShard {
    int rank, world;            // world == 1: unsharded
    int64_t global_offset;      // first sample of this shard in the capture
    int64_t n_total;
    int emit_tail;
};

static int finish_tiles(urh_ctx* ctx, int64_t n, int tol, bool is_ask, uint32_t sps, const UrhTileSummary* tiles, const uint32_t* staging,
                        int stage_cap, const int16_t* d_init, const FinishShard& sh, int64_t* k) {
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const bool sharded = sh.world > 1;
    RunCarry* carry;
    int32_t *head_rel, *prev_cls;
    int64_t *row_off, *prev_fired, *d_small;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &carry));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &head_rel));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &prev_cls));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &row_off));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &prev_fired));
    // small block (int64 units): [0..1] stage D's outputs, [2] -1 (no previous firing), [4..5] RunCarry total, [6..7] CandAgg total,
    // [8..9] FireAgg total, [10..11] folded run carry, [12] folded previous class (int16), [13] folded previous firing,
    // [16..19] stage-1 message, [32..) gathered messages: world x 4 (stage 1), world x 2 (stage 2), world x 2 (stage 3)
    URH_CHECK(urh_arena(ctx, (size_t)(32 + 8 * (sharded ? sh.world : 0)), &d_small));
    URH_CUDA(ctx, cudaMemsetAsync(d_small, 0xff, 4 * sizeof(int64_t), ctx->stream));
    RunCarry* d_tot_run = (RunCarry*)(d_small + 4);
    CandAgg* d_tot_cand = (CandAgg*)(d_small + 6);
    FireAgg* d_tot_fire = (FireAgg*)(d_small + 8);
    RunCarry* d_xcarry = (RunCarry*)(d_small + 10);
    int16_t* d_prev0 = (int16_t*)(d_small + 12);
    int64_t* d_xprev = d_small + 13;
    int64_t* d_msg1 = d_small + 16;
    int64_t* d_all1 = d_small + 32;
    int64_t* d_all2 = d_all1 + 4 * (sharded ? sh.world : 0);
    int64_t* d_all3 = d_all2 + 2 * (sharded ? sh.world : 0);

    RunCarry rc_ident;
    rc_ident.len = 0; rc_ident.cls = 0; rc_ident.flags = 2 | 1;
    ScanRunCarry fa;
    fa.tiles = tiles; fa.n = n; fa.carry = carry;
    URH_CHECK((urhts::scan<RunCarry, RunCarryOp, ScanRunCarry>(ctx, ntiles, rc_ident, RunCarryOp(), fa, d_tot_run)));
    if (sharded) {
        URH_LAUNCH(ctx, k_pack_stage1, 1, 1, 0, d_init, (const RunCarry*)d_tot_run, d_msg1);
        URH_TL_MARK(ctx, "x4 run carry: enter");
        URH_CHECK(urh_coll_allgather(ctx, d_msg1, d_all1, 4 * sizeof(int64_t)));
        URH_TL_MARK(ctx, "x4 run carry: done");
        URH_LAUNCH(ctx, k_fold_carry, 1, 1, 0, (const int64_t*)d_all1, sh.rank, d_xcarry);
    }
    ScanCandidates fb;
    fb.tiles = tiles; fb.staging = staging; fb.stage_cap = stage_cap; fb.carry = carry; fb.xcarry = sharded ? d_xcarry : nullptr;
    fb.tol = tol; fb.head_rel = head_rel; fb.prev_cls = prev_cls;
    CandAgg ca_ident;
    ca_ident.cnt = 0; ca_ident.last_cls = CLS_NONE; ca_ident.pad = 0;
    URH_CHECK((urhts::scan<CandAgg, CandOp, ScanCandidates>(ctx, ntiles, ca_ident, CandOp(), fb, d_tot_cand)));
    const int16_t* prev0 = d_init;
    if (sharded) {
        URH_TL_MARK(ctx, "x5 candidates: enter");
        URH_CHECK(urh_coll_allgather(ctx, d_tot_cand, d_all2, sizeof(CandAgg)));
        URH_TL_MARK(ctx, "x5 candidates: done");
        URH_LAUNCH(ctx, k_fold_prev_cls, 1, 1, 0, (const int64_t*)d_all2, sh.rank, (const int64_t*)d_all1, d_prev0);
        prev0 = d_prev0;
    }
    ScanFirings fc;
    fc.tiles = tiles; fc.staging = staging; fc.stage_cap = stage_cap; fc.head_rel = head_rel; fc.prev_cls = prev_cls; fc.d_prev0 = prev0;
    fc.global_offset = sh.global_offset; fc.row_off = row_off; fc.prev_fired = prev_fired;
    FireAgg fi_ident;
    fi_ident.fired = 0; fi_ident.last_pos = -1;
    URH_CHECK((urhts::scan<FireAgg, FireOp, ScanFirings, 4>(ctx, ntiles, fi_ident, FireOp(), fc, d_tot_fire)));   // heavy load(): thin blocks
    const int64_t* xprev = d_small + 2;
    if (sharded) {
        URH_TL_MARK(ctx, "x6 firings: enter");
        URH_CHECK(urh_coll_allgather(ctx, d_tot_fire, d_all3, sizeof(FireAgg)));
        URH_TL_MARK(ctx, "x6 firings: done");
        URH_LAUNCH(ctx, k_fold_prev_fired, 1, 1, 0, (const int64_t*)d_all3, sh.rank, d_xprev);
        xprev = d_xprev;
    }

    // rows: straight into the pulse buffer (ASK: into scratch, merged afterwards)
    int64_t cap_rows = (int64_t)ctx->pulses_cap_rows;
    const int64_t guess = n / 64 + 1024;
    if (cap_rows < guess) {
        URH_CHECK(urh_ensure_pulses(ctx, (size_t)guess));
        cap_rows = (int64_t)ctx->pulses_cap_rows;
    }
    int64_t* raw = ctx->pulses;
    int64_t raw_cap = cap_rows;
    if (is_ask) URH_CHECK(urh_arena(ctx, (size_t)raw_cap * 2, &raw));
    int64_t got[2] = {0, 0};
    for (int attempt = 0; attempt < 2; attempt++) {
        URH_LAUNCH(ctx, k_finish_rows, (unsigned)urh_div_up(ntiles, 256), 256, 0, tiles, staging, stage_cap, (const int32_t*)head_rel,
                   (const int32_t*)prev_cls, prev0, (const int64_t*)row_off, (const int64_t*)prev_fired, xprev, ntiles, sh.global_offset,
                   sh.n_total, tol, is_ask ? 1 : 0, (int64_t)sps, sh.emit_tail, raw, raw_cap, d_small);
        if (sharded && attempt == 0) URH_TL_MARK(ctx, "rows written");
        URH_CHECK(urh_read_i64(ctx, d_small, 2, got));
        if (sharded) URH_CHECK(urh_p2p_check(ctx));   // a mailbox exchange of this step (or of the center chain before it) timed out?
        if (got[0] <= raw_cap) break;
        if (attempt == 1) URH_FAIL(ctx, URH_ERR_CUDA, "finish_tiles: row buffer overflow after regrowth");
        // more rows than guessed: grow and repeat stage D only
        if (is_ask) {
            URH_CHECK(urh_arena(ctx, (size_t)got[0] * 2, &raw));
            raw_cap = got[0];
        } else {
            URH_CHECK(urh_ensure_pulses(ctx, (size_t)got[0]));
            raw = ctx->pulses;
            raw_cap = (int64_t)ctx->pulses_cap_rows;
        }
    }
    int64_t K = got[0];
    if (is_ask && K > 0) {
        URH_CHECK(urh_ensure_pulses(ctx, (size_t)K));
        URH_CUDA(ctx, cudaMemsetAsync(ctx->pulses, 0, (size_t)K * 2 * sizeof(int64_t), ctx->stream));
        ScanMergeRows fm;
        fm.raw = raw; fm.rows = K; fm.n_total = sh.n_total; fm.has_tail = (sh.emit_tail && K > got[1]) ? 1 : 0; fm.out = ctx->pulses;
        fm.d_k = d_small;
        URH_CHECK((urhts::scan<int64_t, AddI64, ScanMergeRows>(ctx, K, (int64_t)0, AddI64(), fm, (int64_t*)nullptr)));
        URH_CHECK(urh_read_i64(ctx, d_small, 1, &K));
    }
    ctx->pulses_k = K;
    *k = K;
    return URH_OK;
}

int urh_finish_local(urh_ctx* ctx, int64_t n, int tol, bool is_ask, uint32_t sps, const UrhTileSummary* tiles, const uint32_t* staging,
                     int stage_cap, const int16_t* d_init, int64_t* k) {
    FinishShard sh;
    sh.rank = 0; sh.world = 1; sh.global_offset = 0; sh.n_total = n; sh.emit_tail = 1;
    return finish_tiles(ctx, n, tol, is_ask, sps, tiles, staging, stage_cap, d_init, sh, k);
}

// One shard of a capture spread over the ranks of the context's NCCL communicator: same stages, the three scan totals
// exchanged on the stream.  Every rank ends with the rows of its own shard (urh_fetch_pulses); equal states meeting at a
// shard edge are joined by the consumer.
int urh_finish_shard(urh_ctx* ctx, int64_t n, int tol, bool is_ask, uint32_t sps, const UrhTileSummary* tiles, const uint32_t* staging,
                     int stage_cap, const int16_t* d_init, int64_t global_offset, int64_t n_total, int64_t* k) {
    FinishShard sh;
    sh.rank = ctx->nccl_rank; sh.world = ctx->nccl_world; sh.global_offset = global_offset; sh.n_total = n_total;
    sh.emit_tail = (ctx->nccl_rank == ctx->nccl_world - 1) ? 1 : 0;
    return finish_tiles(ctx, n, tol, is_ask, sps, tiles, staging, stage_cap, d_init, sh, k);
}
