// Sparse stages of the digitizer (see sparse.cuh).
#include "sparse.cuh"
#include "scan.cuh"

// ---- run stitching across tiles (RunCarry / RunCarryOp: sparse.cuh) -------------------------------------
__global__ void k_tile_elems(const UrhTileSummary* __restrict__ tiles, int64_t ntiles, int64_t n, RunCarry* __restrict__ e) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int64_t rem = n - t * URH_TILE;
    const int tile_len = rem < URH_TILE ? (int)rem : URH_TILE;
    const UrhTileSummary s = tiles[t];
    RunCarry r;
    r.len = s.tail_len;
    r.cls = s.last_cls;
    r.flags = (s.head_len == tile_len) ? 1 : 0;
    e[t] = r;
}

// head candidate of each tile from the carry of all preceding tiles
__global__ void k_tile_heads(const UrhTileSummary* __restrict__ tiles, const RunCarry* __restrict__ carry,
                             int64_t ntiles, int tol, int32_t* __restrict__ head_rel, int64_t* __restrict__ total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const UrhTileSummary s = tiles[t];
    const RunCarry c = carry[t];
    int64_t start_len = 0;
    if (!(c.flags & 2) && c.cls == s.first_cls) start_len = c.len;
    int32_t rel = -1;
    if (start_len <= tol && (int64_t)tol < start_len + s.head_len) rel = (int32_t)(tol - start_len);
    head_rel[t] = rel;
    total[t] = (int64_t)s.ncand + (rel >= 0 ? 1 : 0);
}

// one warp per tile: head candidate first, then the staged interior candidates
__global__ void k_gather(const UrhTileSummary* __restrict__ tiles, const uint32_t* __restrict__ staging, int stage_cap,
                         const int32_t* __restrict__ head_rel, const int64_t* __restrict__ offset, int64_t ntiles,
                         int64_t global_offset, int64_t* __restrict__ pos, int16_t* __restrict__ cls) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= ntiles) return;
    const UrhTileSummary s = tiles[t];
    int64_t o = offset[t];
    const int64_t base = t * URH_TILE + global_offset;
    const int32_t rel = head_rel[t];
    if (rel >= 0) {
        if (lane == 0) {
            pos[o] = base + rel;
            cls[o] = s.first_cls;
        }
        o++;
    }
    const uint32_t* st = staging + t * (int64_t)stage_cap;
    for (int j = lane; j < s.ncand; j += 32) {
        const uint32_t v = st[j];
        pos[o + j] = base + (v >> 16);
        cls[o + j] = (int16_t)((int)(v & 0xffff) - 1);
    }
}

__global__ void k_apply_carry(RunCarry* __restrict__ carry, int64_t ntiles, RunCarry in) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    carry[t] = RunCarryOp()(in, carry[t]);
}

int urh_shard_run_total(urh_ctx* ctx, int64_t n, const UrhTileSummary* tiles, int64_t* h_out) {
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    RunCarry* carry;
    RunCarry* d_total_run;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &carry));
    URH_CHECK(urh_arena(ctx, 2, &d_total_run));
    URH_LAUNCH(ctx, k_tile_elems, (unsigned)urh_div_up(ntiles, 256), 256, 0, tiles, ntiles, n, carry);
    RunCarry ident;
    ident.len = 0; ident.cls = 0; ident.flags = 2 | 1;
    URH_CHECK((urhscan::device_scan<RunCarry, RunCarryOp>(ctx, carry, ntiles, RunCarryOp(), ident, true, d_total_run)));
    int64_t raw[2];
    URH_CHECK(urh_read_i64(ctx, (const int64_t*)d_total_run, 2, raw));
    RunCarry tr;
    memcpy(&tr, raw, sizeof(tr));
    h_out[0] = tr.cls;
    h_out[1] = tr.len;
    h_out[2] = (tr.flags & 1) ? 1 : 0;
    return URH_OK;
}

int urh_collect_candidates(urh_ctx* ctx, int64_t n, int tol, const UrhTileSummary* tiles, const uint32_t* staging,
                           int stage_cap, UrhCandidates* out) {
    UrhShardCarry none;
    none.valid = 0; none.cls = 0; none.len = 0;
    return urh_collect_candidates_shard(ctx, n, tol, tiles, staging, stage_cap, none, 0, out);
}

int urh_collect_candidates_shard(urh_ctx* ctx, int64_t n, int tol, const UrhTileSummary* tiles, const uint32_t* staging,
                                 int stage_cap, UrhShardCarry carry_in, int64_t global_offset, UrhCandidates* out) {
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    RunCarry* carry;
    int32_t* head_rel;
    int64_t* total;
    int64_t* d_count;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &carry));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &head_rel));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &total));
    URH_CHECK(urh_arena(ctx, 4, &d_count));
    const unsigned g = (unsigned)urh_div_up(ntiles, 256);
    URH_LAUNCH(ctx, k_tile_elems, g, 256, 0, tiles, ntiles, n, carry);
    RunCarry ident;
    ident.len = 0;
    ident.cls = 0;
    ident.flags = 2 | 1;
    RunCarry* d_total_run;
    URH_CHECK(urh_arena(ctx, 2, &d_total_run));
    URH_CHECK((urhscan::device_scan<RunCarry, RunCarryOp>(ctx, carry, ntiles, RunCarryOp(), ident, true, d_total_run)));
    if (carry_in.valid) {
        RunCarry in;
        in.len = carry_in.len; in.cls = carry_in.cls; in.flags = 0;
        URH_LAUNCH(ctx, k_apply_carry, g, 256, 0, carry, ntiles, in);
    }
    URH_LAUNCH(ctx, k_tile_heads, g, 256, 0, tiles, carry, ntiles, tol, head_rel, total);
    URH_CHECK((urhscan::device_scan<int64_t, urhscan::AddI64>(ctx, total, ntiles, urhscan::AddI64(), (int64_t)0, true, d_count)));
    int64_t C = 0;
    URH_CHECK(urh_read_i64(ctx, d_count, 1, &C));
    {
        int64_t raw[2];
        URH_CHECK(urh_read_i64(ctx, (const int64_t*)d_total_run, 2, raw));
        RunCarry tr;
        memcpy(&tr, raw, sizeof(tr));
        out->last_cls = tr.cls;
        out->last_len = tr.len;
    }
    out->count = C;
    out->pos = nullptr;
    out->cls = nullptr;
    if (C == 0) return URH_OK;
    URH_CHECK(urh_arena(ctx, (size_t)C, &out->pos));
    URH_CHECK(urh_arena(ctx, (size_t)C, &out->cls));
    const unsigned gg = (unsigned)urh_div_up(ntiles * 32, 256);
    URH_LAUNCH(ctx, k_gather, gg, 256, 0, tiles, staging, stage_cap, head_rel, total, ntiles, global_offset, out->pos, out->cls);
    return URH_OK;
}

// ---- grab_pulse_lens on the candidate table ---------------------------------------------------------
__global__ void k_fire_flags(const int16_t* __restrict__ cls, const int16_t* __restrict__ init, int64_t C,
                             int64_t* __restrict__ fire) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const int prev = j ? cls[j - 1] : *init;
    fire[j] = (cls[j] != prev) ? 1 : 0;
}

__global__ void k_fired_rows(const int64_t* __restrict__ pos, const int16_t* __restrict__ cls,
                             const int16_t* __restrict__ init, const int64_t* __restrict__ fidx, int64_t C,
                             int64_t* __restrict__ fpos, int64_t* __restrict__ st) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const int prev = j ? cls[j - 1] : *init;
    if (cls[j] != prev) {
        const int64_t f = fidx[j];
        fpos[f] = pos[j];
        st[f] = prev;  // the state that ends here (cur_state at pyx:475-479)
    }
}

// pulse lengths (pyx:476-482), ASK short-pause relabel (pyx:471-473), tail row (pyx:485-493).
// Sharded captures: prev_fired (position of the last firing in the preceding shards, -1 if none) replaces the
// "first pulse" rule, and only the last shard emits the tail row.
__global__ void k_row_lengths(const int64_t* __restrict__ fpos, int64_t* __restrict__ st, int64_t* __restrict__ ln,
                              int64_t F, int64_t n, int tol, int is_ask, int64_t sps, const int16_t* __restrict__ cls,
                              int64_t C, const int16_t* __restrict__ init, int64_t prev_fired, int emit_tail) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f > F) return;
    if (f == F) {
        if (!emit_tail) return;
        st[f] = C ? cls[C - 1] : *init;
        ln[f] = F ? (n - 1 - fpos[F - 1]) : (prev_fired >= 0 ? (n - 1 - prev_fired) : (n - tol));
        return;
    }
    const int64_t rec = f ? (fpos[f] - fpos[f - 1]) : (prev_fired >= 0 ? (fpos[0] - prev_fired) : (fpos[0] + 1 - tol));
    if (is_ask && st[f] == -1 && rec < sps) st[f] = 0;
    ln[f] = rec;
}

__global__ void k_row_heads(const int64_t* __restrict__ st, int64_t rows, int64_t* __restrict__ head) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    head[r] = (r == 0 || st[r] != st[r - 1]) ? 1 : 0;
}

__global__ void k_row_merge(const int64_t* __restrict__ st, const int64_t* __restrict__ ln,
                            const int64_t* __restrict__ seg_incl, int64_t rows, int64_t* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int64_t o = seg_incl[r] - 1;
    if (r == 0 || st[r] != st[r - 1]) out[2 * o] = st[r];
    atomicAdd((unsigned long long*)&out[2 * o + 1], (unsigned long long)ln[r]);
}

int urh_fire_stage(urh_ctx* ctx, const UrhCandidates& cand, const int16_t* d_prev_cls, UrhFireState* fs, int64_t* last_fired_pos) {
    const int64_t C = cand.count;
    fs->C = C;
    fs->F = 0;
    fs->pos = cand.pos;
    fs->cls = cand.cls;
    fs->d_prev_cls = d_prev_cls;
    fs->fire = nullptr;
    int64_t* d_tot = nullptr;
    URH_CHECK(urh_arena(ctx, 4, &d_tot));
    if (C > 0) {
        URH_CHECK(urh_arena(ctx, (size_t)C, &fs->fire));
        URH_LAUNCH(ctx, k_fire_flags, (unsigned)urh_div_up(C, 256), 256, 0, cand.cls, d_prev_cls, C, fs->fire);
        URH_CHECK((urhscan::device_scan<int64_t, urhscan::AddI64>(ctx, fs->fire, C, urhscan::AddI64(), (int64_t)0, true, d_tot)));
        URH_CHECK(urh_read_i64(ctx, d_tot, 1, &fs->F));
    }
    const int64_t F = fs->F;
    URH_CHECK(urh_arena(ctx, (size_t)F + 1, &fs->fpos));
    URH_CHECK(urh_arena(ctx, (size_t)F + 1, &fs->st));
    URH_CHECK(urh_arena(ctx, (size_t)F + 1, &fs->ln));
    URH_CHECK(urh_arena(ctx, (size_t)F + 1, &fs->head));
    if (F > 0) URH_LAUNCH(ctx, k_fired_rows, (unsigned)urh_div_up(C, 256), 256, 0, cand.pos, cand.cls, d_prev_cls, fs->fire, C, fs->fpos, fs->st);
    if (last_fired_pos) {
        *last_fired_pos = -1;
        if (F > 0) URH_CHECK(urh_read_i64(ctx, fs->fpos + F - 1, 1, last_fired_pos));
    }
    return URH_OK;
}

int urh_rows_stage(urh_ctx* ctx, const UrhFireState& fs, int64_t n, int tol, bool is_ask, uint32_t sps, int64_t prev_fired,
                   bool emit_tail, int64_t* k) {
    const int64_t F = fs.F, C = fs.C;
    int64_t *fpos = fs.fpos, *st = fs.st, *ln = fs.ln, *head = fs.head;
    const int64_t cand_rows = emit_tail ? F + 1 : F;
    if (cand_rows == 0) {
        ctx->pulses_k = 0;
        *k = 0;
        return URH_OK;
    }
    URH_LAUNCH(ctx, k_row_lengths, (unsigned)urh_div_up(F + 1, 256), 256, 0, fpos, st, ln, F, n, tol, is_ask ? 1 : 0,
               (int64_t)sps, fs.cls, C, fs.d_prev_cls, prev_fired, emit_tail ? 1 : 0);
    URH_LAUNCH(ctx, k_row_heads, (unsigned)urh_div_up(cand_rows, 256), 256, 0, st, cand_rows, head);
    URH_CHECK((urhscan::device_scan<int64_t, urhscan::AddI64>(ctx, head, cand_rows, urhscan::AddI64(), (int64_t)0, false, nullptr)));
    // merged row count before/with the tail row
    int64_t hv[2] = {0, 0};
    if (emit_tail) {
        if (F > 0) {
            URH_CHECK(urh_read_i64(ctx, head + F - 1, 2, hv));
        } else {
            URH_CHECK(urh_read_i64(ctx, head, 1, &hv[1]));
        }
    } else {
        URH_CHECK(urh_read_i64(ctx, head + F - 1, 1, &hv[0]));
    }
    const int64_t merged_before_tail = F > 0 ? hv[0] : 0;
    // pyx:487: the tail row is only appended while cur_index < len(result) == n
    const bool keep_tail = emit_tail && merged_before_tail < n;
    const int64_t rows = keep_tail ? F + 1 : F;
    const int64_t K = keep_tail ? hv[1] : merged_before_tail;
    URH_CHECK(urh_ensure_pulses(ctx, (size_t)K));
    URH_CUDA(ctx, cudaMemsetAsync(ctx->pulses, 0, (size_t)K * 2 * sizeof(int64_t), ctx->stream));
    if (rows > 0) URH_LAUNCH(ctx, k_row_merge, (unsigned)urh_div_up(rows, 256), 256, 0, st, ln, head, rows, ctx->pulses);
    ctx->pulses_k = K;
    *k = K;
    return URH_OK;
}

int urh_pulses_from_candidates(urh_ctx* ctx, int64_t n, int tol, bool is_ask, uint32_t sps, const UrhCandidates& cand,
                               const int16_t* d_init_cls, int64_t* k) {
    UrhFireState fs;
    URH_CHECK(urh_fire_stage(ctx, cand, d_init_cls, &fs, nullptr));
    return urh_rows_stage(ctx, fs, n, tol, is_ask, sps, -1, true, k);
}
