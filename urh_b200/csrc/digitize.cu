// afp_demod (ASK/FSK), grab_pulse_lens and the fused demod+digitize path.
//
// Reference: src/urh/cythonext/signal_functions.pyx:333-378 (afp_demod), :392-495 (grab_pulse_lens).
// See dense.cuh for the dense pass and DESIGN.md for the run/candidate restatement of the digitizer.
#include "dense.cuh"
#include "scan.cuh"
#include "sparse.cuh"
#ifndef URH_FAST_MIN_BLOCKS
#define URH_FAST_MIN_BLOCKS 4
#endif
#include "fsk_fast.cuh"
#include "dense_f32.cuh"

#include <math.h>
#include <stdlib.h>

// =====================================================================================================
// Dense kernels
// =====================================================================================================

// IQ source: demodulate (+ optionally write qad, + optionally digitize).
template <int DT, int MOD, bool DIGITIZE>
__global__ void __launch_bounds__(URH_WARPS_PER_BLOCK * 32)
k_dense_iq(const void* __restrict__ iq, int64_t n, UrhDemodParams dp, float* __restrict__ qad_out, int vec_in,
           int vec_out, const __grid_constant__ UrhClassify cls, int tol, UrhTileSummary* __restrict__ tiles,
           uint32_t* __restrict__ staging, int stage_cap, int16_t* __restrict__ init_cls, int cls_of_zero,
           int64_t tile_begin, int64_t tile_count, int has_halo, UrhTileStats* __restrict__ tile_stats) {
    const int lane = threadIdx.x & 31;
    const int64_t tile_rel = (int64_t)blockIdx.x * URH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (tile_rel >= tile_count) return;
    const int64_t tile = tile_begin + tile_rel;
    const int64_t tile_start = tile * URH_TILE;
    if (tile_start >= n) return;
    const int tile_len = (int)((n - tile_start) < URH_TILE ? (n - tile_start) : URH_TILE);
    const int iters = (tile_len + 63) >> 6;

    UrhRunTracker rt;
    if (DIGITIZE) rt.init(tol, staging + tile * (int64_t)stage_cap);
    UrhStatAcc acc;
    acc.init();

    // FSK: (A, B) terms of the sample preceding the tile's first sample
    float cA = 0.0f, cB = 0.0f;
    // (a shard of a larger capture has its predecessor sample stored right before iq: has_halo)
    if (MOD == URH_MOD_FSK && (tile_start > 0 || has_halo) && lane == 0) {
        typedef typename UrhElem<DT>::type E;
        const E* pp = (const E*)iq + 2 * (tile_start - 1);
        const UrhFskTerms t = urh_fsk_terms((float)pp[0], (float)pp[1]);
        cA = t.A; cB = t.B;
    }

    UrhPair cur = urh_load_pair<DT>(iq, tile_start + 2 * lane, n, vec_in != 0);
    for (int it = 0; it < iters; it++) {
        const int64_t pos0 = tile_start + (int64_t)it * 64 + 2 * lane;
        UrhPair nxt;
        if (it + 1 < iters) nxt = urh_load_pair<DT>(iq, pos0 + 64, n, vec_in != 0);

        // noise gate: magnitude = re*re + im*im <= noise_sqrd (pyx:366-369)
        const float m0 = __fadd_rn(__fmul_rn(cur.r0, cur.r0), __fmul_rn(cur.i0, cur.i0));
        const float m1 = __fadd_rn(__fmul_rn(cur.r1, cur.r1), __fmul_rn(cur.i1, cur.i1));
        float s0 = dp.noise_value, s1 = dp.noise_value;
        if (MOD == URH_MOD_FSK) {
            const UrhFskTerms t0 = urh_fsk_terms(cur.r0, cur.i0);
            const UrhFskTerms t1 = urh_fsk_terms(cur.r1, cur.i1);
            float pA = __shfl_up_sync(URH_FULL_MASK, t1.A, 1);
            float pB = __shfl_up_sync(URH_FULL_MASK, t1.B, 1);
            if (lane == 0) { pA = cA; pB = cB; }
            cA = __shfl_sync(URH_FULL_MASK, t1.A, 31);
            cB = __shfl_sync(URH_FULL_MASK, t1.B, 31);
            if (!(m0 <= dp.noise_sqrd)) s0 = urh_fsk_angle(pA, pB, t0.C, t0.D);
            if (!(m1 <= dp.noise_sqrd)) s1 = urh_fsk_angle(t0.A, t0.B, t1.C, t1.D);
        } else if (MOD == URH_MOD_ASK) {
            if (!(m0 <= dp.noise_sqrd)) s0 = __fdiv_rn(__fsqrt_rn(m0), dp.max_mag);
            if (!(m1 <= dp.noise_sqrd)) s1 = __fdiv_rn(__fsqrt_rn(m1), dp.max_mag);
        }
        if (pos0 == 0 && !has_halo) s0 = dp.noise_value;  // result[0] = NOISE (pyx:361); shards: only the capture's first sample

        const bool v0 = pos0 < n, v1 = pos0 + 1 < n;
        if (tile_stats) {
            if (v0) { acc.add(s0); acc.all_noise = acc.all_noise && (s0 == dp.noise_value); }
            if (v1) { acc.add(s1); acc.all_noise = acc.all_noise && (s1 == dp.noise_value); }
        }
        if (qad_out) {
            if (v1 && vec_out) urh_stg_f2(qad_out + pos0, s0, s1);
            else {
                if (v0) qad_out[pos0] = s0;
                if (v1) qad_out[pos0 + 1] = s1;
            }
        }
        if (DIGITIZE) {
            const int c0 = urh_classify(s0, cls), c1 = urh_classify(s1, cls);
            if (pos0 == 0) *init_cls = (int16_t)((s0 == cls.noise_value) ? -1 : cls_of_zero);
            rt.feed(it, c0, c1, v0, v1, lane);
        }
        cur = nxt;
    }
    if (tile_stats) acc.store(tile_stats + tile, lane);
    if (DIGITIZE) rt.finish(tile_len, tiles + tile, lane);
}

// Fast kernel: full, aligned, order-2 FSK tiles [tile_begin, tile_begin + tile_count), tile_begin >= 1
// (fsk_fast.cuh: packed f32x2 math, same bits as the generic kernel).
template <int DT, bool DIGITIZE, bool WRITE, bool STATS>
__global__ void __launch_bounds__(URH_WARPS_PER_BLOCK * 32, URH_FAST_MIN_BLOCKS)
k_fsk_fast(const void* __restrict__ iq, int64_t n, UrhDemodParams dp, float* __restrict__ qad_out, float thr0,
           float cls_noise, int tol, UrhTileSummary* __restrict__ tiles, uint32_t* __restrict__ staging, int stage_cap,
           int64_t tile_begin, int64_t tile_count, UrhTileStats* __restrict__ tile_stats) {
    const int lane = threadIdx.x & 31;
    const int64_t tile_rel = (int64_t)blockIdx.x * URH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (tile_rel >= tile_count) return;
    const int64_t tile = tile_begin + tile_rel;
    UrhRunTracker rt;
    if (DIGITIZE) rt.init(tol, staging + tile * (int64_t)stage_cap);
    UrhOne one;
    one.p = dp.one;
    one.m = dp.mone;
    urh_fsk_full_tile<DT, DIGITIZE, WRITE, STATS>(iq, n, tile * URH_TILE, dp, qad_out, thr0, cls_noise, rt, lane, one,
                                                  STATS ? tile_stats + tile : nullptr, 0u, DIGITIZE ? tiles + tile : nullptr);
}

// The same kernel with the input staged through the shared-memory FIFO.
#define URH_FSK_FIFO 3
template <int DT, bool DIGITIZE, bool WRITE, bool STATS>
__global__ void __launch_bounds__(URH_WARPS_PER_BLOCK * 32, (DT == URH_DT_F32) ? 5 : 4)   // (the integer variants spill at 48 registers)
k_fsk_fifo(const void* __restrict__ iq, int64_t n, UrhDemodParams dp, float* __restrict__ qad_out, float thr0,
           float cls_noise, int tol, UrhTileSummary* __restrict__ tiles, uint32_t* __restrict__ staging, int stage_cap,
           int64_t tile_begin, int64_t tile_count, UrhTileStats* __restrict__ tile_stats) {
    __shared__ __align__(16) unsigned char s_fifo[URH_WARPS_PER_BLOCK][URH_FSK_FIFO + 1][64 * 2 * sizeof(typename UrhElem<DT>::type)];
    const int lane = threadIdx.x & 31;
    const int64_t tile_rel = (int64_t)blockIdx.x * URH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (tile_rel >= tile_count) return;
    const int64_t tile = tile_begin + tile_rel;
    UrhRunTracker rt;
    if (DIGITIZE) rt.init(tol, staging + tile * (int64_t)stage_cap);
    UrhOne one;
    one.p = dp.one;
    one.m = dp.mone;
    const uint32_t fifo = (uint32_t)__cvta_generic_to_shared(&s_fifo[threadIdx.x >> 5][0][0]);
    urh_fsk_full_tile<DT, DIGITIZE, WRITE, STATS, URH_FSK_FIFO>(iq, n, tile * URH_TILE, dp, qad_out, thr0, cls_noise, rt, lane, one,
                                                                        STATS ? tile_stats + tile : nullptr, fifo,
                                                                        DIGITIZE ? tiles + tile : nullptr);
}

// =====================================================================================================
// Host side
// =====================================================================================================

extern "C" int urh_get_center_thresholds(float center, float spacing, int modulation_order, float* h_out) {
    // signal_functions.pyx:380-390 — float32 arithmetic on (int * float)
    const int n = modulation_order / 2;
    for (int i = 0; i < n; i++) {
        volatile float t = (float)(n - (i + 1)) * spacing;
        h_out[i] = center - t;
    }
    for (int i = n; i < modulation_order - 1; i++) {
        volatile float t = (float)(i + 1 - n) * spacing;
        h_out[i] = center + t;
    }
    return URH_OK;
}

static int fill_classify(urh_ctx* ctx, UrhClassify* C, int mod_type, float center, uint8_t bits_per_symbol,
                         float spacing) {
    if (bits_per_symbol > 8) URH_FAIL(ctx, URH_ERR_INVALID, "bits_per_symbol %d > 8 not supported", (int)bits_per_symbol);
    memset(C, 0, sizeof(*C));
    C->noise_value = urh_noise_value(mod_type);
    C->order = 1 << bits_per_symbol;
    if (C->order > 1) urh_get_center_thresholds(center, spacing, C->order, C->thr);
    return URH_OK;
}

static int host_classify(float s, const UrhClassify& C) {
    int c = C.order - 1;
    for (int k = 0; k < C.order - 1; k++)
        if (s <= C.thr[k]) { c = k; break; }
    return c;
}

static float max_magnitude_for(int dtype) {
    // signal_functions.pyx:343-352: double sqrt of an integer constant, stored in a float
    switch (dtype) {
        case URH_DT_I8: return (float)sqrt(127.0 * 127.0 + 128.0 * 128.0);
        case URH_DT_U8: return (float)sqrt(255.0 * 255.0);
        case URH_DT_I16: return (float)sqrt(32768.0 * 32768.0 + 32767.0 * 32767.0);
        case URH_DT_U16: return (float)sqrt(65535.0 * 65535.0);
        default: return (float)sqrt(2.0);
    }
}

static bool iq_vec_aligned(const void* p, int dtype) { return ((uintptr_t)p % (2 * (size_t)urh_iq_bytes(dtype))) == 0; }

// tile_lo / tile_hi: restrict the pass to tiles [tile_lo, tile_hi) (chunked ingest: a chunk is demodulated as soon as its
// upload has landed); tile_hi < 0 = all tiles.
template <int DT, int MOD, bool DIG>
static int launch_dense_iq_t(urh_ctx* ctx, const void* d_iq, int64_t n, const UrhDemodParams& dp, float* d_qad,
                             const UrhClassify& cls, int tol, UrhTileSummary* tiles, uint32_t* staging,
                             int stage_cap, int16_t* init_cls, int cls_of_zero, int has_halo = 0,
                             UrhTileStats* tile_stats = nullptr, int64_t tile_lo = 0, int64_t tile_hi = -1) {
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    if (tile_hi < 0 || tile_hi > ntiles) tile_hi = ntiles;
    const int vec_in = iq_vec_aligned(d_iq, DT) ? 1 : 0;
    const int vec_out = (d_qad && ((uintptr_t)d_qad % 8) == 0) ? 1 : 0;
    const int threads = URH_WARPS_PER_BLOCK * 32;
    auto generic = [&](int64_t begin, int64_t end) -> int {   // tiles [begin, end) clipped to the requested range
        begin = begin < tile_lo ? tile_lo : begin;
        end = end > tile_hi ? tile_hi : end;
        const int64_t count = end - begin;
        if (count <= 0) return URH_OK;
        URH_LAUNCH(ctx, (k_dense_iq<DT, MOD, DIG>), (unsigned)urh_div_up(count, URH_WARPS_PER_BLOCK), threads, 0, d_iq, n, dp,
                   d_qad, vec_in, vec_out, cls, tol, tiles, staging, stage_cap, init_cls, cls_of_zero, begin, count, has_halo,
                   tile_stats);
        return URH_OK;
    };
    // FSK on aligned buffers with a binary digitizer: tiles 1 .. nfull-1 take the packed-f32x2 kernel
    const int64_t nfull = n / URH_TILE;
    const bool fast = MOD == URH_MOD_FSK && vec_in && (!d_qad || vec_out) && (!DIG || cls.order == 2) && nfull > 1;
    URH_PROF_BEGIN(ctx);
    if (fast) {
        const int64_t fb = tile_lo > 1 ? tile_lo : 1, fe = tile_hi < nfull ? tile_hi : nfull;
        if (fe > fb) {
            const unsigned grid = (unsigned)urh_div_up(fe - fb, URH_WARPS_PER_BLOCK);
            // the variant that stages its input through the shared-memory FIFO (float32: 2230 vs 2417 us at 2^30 samples)
            static const bool fifo = getenv("URH_B200_FSK_NO_FIFO") == nullptr;
            const bool ff = fifo;
            if (tile_stats && d_qad && !DIG) {
                if (ff) URH_LAUNCH(ctx, (k_fsk_fifo<DT, false, true, true>), grid, threads, 0, d_iq, n, dp, d_qad, cls.thr[0], cls.noise_value, tol,
                                   tiles, staging, stage_cap, fb, fe - fb, tile_stats);
                else URH_LAUNCH(ctx, (k_fsk_fast<DT, false, true, true>), grid, threads, 0, d_iq, n, dp, d_qad, cls.thr[0], cls.noise_value,
                                tol, tiles, staging, stage_cap, fb, fe - fb, tile_stats);
            } else if (d_qad) {
                if (ff) URH_LAUNCH(ctx, (k_fsk_fifo<DT, DIG, true, false>), grid, threads, 0, d_iq, n, dp, d_qad, cls.thr[0], cls.noise_value, tol,
                                   tiles, staging, stage_cap, fb, fe - fb, nullptr);
                else URH_LAUNCH(ctx, (k_fsk_fast<DT, DIG, true, false>), grid, threads, 0, d_iq, n, dp, d_qad, cls.thr[0], cls.noise_value, tol,
                                tiles, staging, stage_cap, fb, fe - fb, nullptr);
            } else {
                if (ff) URH_LAUNCH(ctx, (k_fsk_fifo<DT, DIG, false, false>), grid, threads, 0, d_iq, n, dp, d_qad, cls.thr[0], cls.noise_value, tol,
                                   tiles, staging, stage_cap, fb, fe - fb, nullptr);
                else URH_LAUNCH(ctx, (k_fsk_fast<DT, DIG, false, false>), grid, threads, 0, d_iq, n, dp, d_qad, cls.thr[0], cls.noise_value, tol,
                                tiles, staging, stage_cap, fb, fe - fb, nullptr);
            }
        }
        URH_CHECK(generic(0, 1));
        URH_CHECK(generic(nfull, ntiles));
    } else {
        URH_CHECK(generic(0, ntiles));
    }
    URH_PROF_END(ctx);
    return URH_OK;
}

template <int MOD, bool DIG>
static int launch_dense_iq_m(urh_ctx* ctx, int dtype, const void* d_iq, int64_t n, const UrhDemodParams& dp,
                             float* d_qad, const UrhClassify& cls, int tol, UrhTileSummary* tiles,
                             uint32_t* staging, int stage_cap, int16_t* init_cls, int cls_of_zero, int has_halo = 0,
                             UrhTileStats* tile_stats = nullptr, int64_t tile_lo = 0, int64_t tile_hi = -1) {
    switch (dtype) {
        case URH_DT_I8: return launch_dense_iq_t<URH_DT_I8, MOD, DIG>(ctx, d_iq, n, dp, d_qad, cls, tol, tiles, staging, stage_cap, init_cls, cls_of_zero, has_halo, tile_stats, tile_lo, tile_hi);
        case URH_DT_U8: return launch_dense_iq_t<URH_DT_U8, MOD, DIG>(ctx, d_iq, n, dp, d_qad, cls, tol, tiles, staging, stage_cap, init_cls, cls_of_zero, has_halo, tile_stats, tile_lo, tile_hi);
        case URH_DT_I16: return launch_dense_iq_t<URH_DT_I16, MOD, DIG>(ctx, d_iq, n, dp, d_qad, cls, tol, tiles, staging, stage_cap, init_cls, cls_of_zero, has_halo, tile_stats, tile_lo, tile_hi);
        case URH_DT_U16: return launch_dense_iq_t<URH_DT_U16, MOD, DIG>(ctx, d_iq, n, dp, d_qad, cls, tol, tiles, staging, stage_cap, init_cls, cls_of_zero, has_halo, tile_stats, tile_lo, tile_hi);
        case URH_DT_F32: return launch_dense_iq_t<URH_DT_F32, MOD, DIG>(ctx, d_iq, n, dp, d_qad, cls, tol, tiles, staging, stage_cap, init_cls, cls_of_zero, has_halo, tile_stats, tile_lo, tile_hi);
        default: URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");
    }
}

int urh_costas_demod(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, float noise_sqrd, int loop_order,
                     float bandwidth, float* d_out);  // costas.cu

static UrhDemodParams make_demod_params(float noise_mag, int mod_type, int dtype) {
    UrhDemodParams dp;
    volatile float nm = noise_mag;
    volatile float sq = nm * nm;
    dp.noise_sqrd = sq;
    dp.noise_value = urh_noise_value(mod_type);
    dp.max_mag = max_magnitude_for(dtype);
    dp.one = 1.0f;
    dp.mone = -1.0f;
    return dp;
}

extern "C" int urh_afp_demod(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, float noise_mag, int mod_type,
                             int mod_order, float costas_loop_bandwidth, float* d_out) {
    if (n < 0) URH_FAIL(ctx, URH_ERR_INVALID, "negative length");
    if (urh_iq_bytes(dtype) == 0) URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");
    if (n == 0) return URH_OK;
    if (n <= 2 || (mod_type != URH_MOD_ASK && mod_type != URH_MOD_FSK && mod_type != URH_MOD_PSK)) {
        // pyx:335-336 (short input) and pyx:360,371-376 (unknown mod_type leaves zeros, result[0] = NOISE)
        URH_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)n * sizeof(float), ctx->stream));
        if (n > 2) {
            const float nv = urh_noise_value(mod_type);
            URH_CUDA(ctx, cudaMemcpyAsync(d_out, &nv, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        return URH_OK;
    }
    const UrhDemodParams dp = make_demod_params(noise_mag, mod_type, dtype);
    if (mod_type == URH_MOD_PSK) return urh_costas_demod(ctx, d_iq, dtype, n, dp.noise_sqrd, mod_order, costas_loop_bandwidth, d_out);
    UrhClassify cls;
    memset(&cls, 0, sizeof(cls));
    if (mod_type == URH_MOD_ASK)
        return launch_dense_iq_m<URH_MOD_ASK, false>(ctx, dtype, d_iq, n, dp, d_out, cls, 0, nullptr, nullptr, 0, nullptr, 0);
    return launch_dense_iq_m<URH_MOD_FSK, false>(ctx, dtype, d_iq, n, dp, d_out, cls, 0, nullptr, nullptr, 0, nullptr, 0);
}

int urh_center_tiles_begin(urh_ctx* ctx, const float* d_x, const UrhTileStats* ts, int64_t n, int64_t* h_total);  // center.cu

// afp_demod (ASK / FSK) that also collects, in the same pass over the IQ samples, what detect_center needs: per-tile
// {count, min, max, sum, sumsq} of the samples it keeps (> -4).  *h_kept = number of kept samples.  The table stays in
// the ctx arena for urh_center_window_stats / urh_center_histogram_tiles (any other ctx call invalidates it).
// halo = 1: the sample preceding d_iq (the previous shard's last) is readable right before it, so qad[0] is a real value
// instead of the capture-start NOISE sentinel.
extern "C" int urh_afp_demod_tiles(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, float noise_mag, int mod_type,
                                   float* d_qad_out, int halo, int64_t* h_kept) {
    if (n <= 2 || (mod_type != URH_MOD_ASK && mod_type != URH_MOD_FSK) || !d_qad_out)
        URH_FAIL(ctx, URH_ERR_INVALID, "urh_afp_demod_tiles: ASK/FSK, n > 2 and a qad buffer are required");
    if (urh_iq_bytes(dtype) == 0) URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");
    urh_arena_reset(ctx);
    const UrhDemodParams dp = make_demod_params(noise_mag, mod_type, dtype);
    UrhClassify cls;
    memset(&cls, 0, sizeof(cls));
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    UrhTileStats* ts;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &ts));
    if (mod_type == URH_MOD_ASK)
        URH_CHECK((launch_dense_iq_m<URH_MOD_ASK, false>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, 0, nullptr, nullptr, 0, nullptr, 0, halo, ts)));
    else
        URH_CHECK((launch_dense_iq_m<URH_MOD_FSK, false>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, 0, nullptr, nullptr, 0, nullptr, 0, halo, ts)));
    return urh_center_tiles_begin(ctx, d_qad_out, ts, n, h_kept);
}

// Shared tail of the two digitizer entry points: tile table + staging -> merged (state, length) rows.
int urh_finish_local(urh_ctx* ctx, int64_t n, int tol, bool is_ask, uint32_t sps, const UrhTileSummary* tiles, const uint32_t* staging,
                     int stage_cap, const int16_t* d_init, int64_t* k);   // finish.cu
int urh_finish_shard(urh_ctx* ctx, int64_t n, int tol, bool is_ask, uint32_t sps, const UrhTileSummary* tiles, const uint32_t* staging,
                     int stage_cap, const int16_t* d_init, int64_t global_offset, int64_t n_total, int64_t* k);   // finish.cu

static int digitize_finish(urh_ctx* ctx, int64_t n, int tol, bool is_ask, uint32_t sps, UrhTileSummary* tiles,
                           uint32_t* staging, int stage_cap, int16_t* d_init, int64_t* k) {
    if (getenv("URH_B200_OLD_FINISH")) {   // the candidate-table formulation (kept for the segmenter; A/B switch for measurements)
        UrhCandidates cand;
        URH_CHECK(urh_collect_candidates(ctx, n, tol, tiles, staging, stage_cap, &cand));
        return urh_pulses_from_candidates(ctx, n, tol, is_ask, sps, cand, d_init, k);
    }
    return urh_finish_local(ctx, n, tol, is_ask, sps, tiles, staging, stage_cap, d_init, k);
}

static int stage_cap_for(int tol) { return URH_TILE / (tol + 1) + 2; }

extern "C" int urh_grab_pulse_lens(urh_ctx* ctx, const float* d_qad, int64_t n, float center, uint16_t tolerance,
                                   int mod_type, uint32_t samples_per_symbol, uint8_t bits_per_symbol,
                                   float center_spacing, int64_t* k) {
    if (!k) return URH_ERR_INVALID;
    *k = 0;
    ctx->pulses_k = 0;
    if (n < 0) URH_FAIL(ctx, URH_ERR_INVALID, "negative length");
    if (n == 0) return URH_OK;  // pyx:416-417 -> empty (0,2) table
    urh_arena_reset(ctx);
    UrhClassify cls;
    URH_CHECK(fill_classify(ctx, &cls, mod_type, center, bits_per_symbol, center_spacing));
    const int tol = tolerance;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = stage_cap_for(tol);
    UrhTileSummary* tiles;
    uint32_t* staging;
    int16_t* d_init;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    URH_CHECK(urh_arena(ctx, 8, &d_init));
    const unsigned grid = (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK);
    const int vec_in = (((uintptr_t)d_qad % 8) == 0) ? 1 : 0;
    URH_PROF_BEGIN(ctx);
    if (cls.order == 2)
        URH_LAUNCH(ctx, (k_dense_f32<SrcQad2, float>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_qad, n, vec_in, cls, tol, tiles,
                   staging, cap, d_init, host_classify(0.0f, cls));
    else
        URH_LAUNCH(ctx, (k_dense_f32<SrcQad, float>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_qad, n, vec_in, cls, tol, tiles,
                   staging, cap, d_init, host_classify(0.0f, cls));
    URH_PROF_END(ctx);
    return digitize_finish(ctx, n, tol, mod_type == URH_MOD_ASK, samples_per_symbol, tiles, staging, cap, d_init, k);
}

extern "C" int urh_demod_digitize(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, float noise_mag,
                                  int mod_type, float center, uint16_t tolerance, uint32_t samples_per_symbol,
                                  uint8_t bits_per_symbol, float center_spacing, float* d_qad_out, int64_t* k) {
    if (!k) return URH_ERR_INVALID;
    *k = 0;
    ctx->pulses_k = 0;
    if (n < 0) URH_FAIL(ctx, URH_ERR_INVALID, "negative length");
    if (urh_iq_bytes(dtype) == 0) URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");
    if (n == 0) return URH_OK;
    if (n <= 2 || (mod_type != URH_MOD_ASK && mod_type != URH_MOD_FSK)) {
        // not a fusable case: run the two reference steps back to back (PSK is a serial recurrence)
        float* q = d_qad_out;
        if (!q) URH_CUDA(ctx, cudaMalloc((void**)&q, (size_t)n * sizeof(float)));
        int rc = urh_afp_demod(ctx, d_iq, dtype, n, noise_mag, mod_type, 1 << bits_per_symbol, 0.1f, q);
        if (rc == URH_OK)
            rc = urh_grab_pulse_lens(ctx, q, n, center, tolerance, mod_type, samples_per_symbol, bits_per_symbol,
                                     center_spacing, k);
        if (!d_qad_out) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(q);
        }
        return rc;
    }
    urh_arena_reset(ctx);
    UrhClassify cls;
    URH_CHECK(fill_classify(ctx, &cls, mod_type, center, bits_per_symbol, center_spacing));
    const UrhDemodParams dp = make_demod_params(noise_mag, mod_type, dtype);
    const int tol = tolerance;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = stage_cap_for(tol);
    UrhTileSummary* tiles;
    uint32_t* staging;
    int16_t* d_init;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    URH_CHECK(urh_arena(ctx, 8, &d_init));
    const int c0 = host_classify(0.0f, cls);
    if (mod_type == URH_MOD_ASK)
        URH_CHECK((launch_dense_iq_m<URH_MOD_ASK, true>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, tol, tiles, staging, cap, d_init, c0)));
    else
        URH_CHECK((launch_dense_iq_m<URH_MOD_FSK, true>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, tol, tiles, staging, cap, d_init, c0)));
    return digitize_finish(ctx, n, tol, mod_type == URH_MOD_ASK, samples_per_symbol, tiles, staging, cap, d_init, k);
}

// =====================================================================================================
// Sharded captures: one contiguous sample range per GPU, 1-sample halo for the FSK conjugate
// product, run-carry descriptors exchanged between ranks, candidate tables gathered to one rank.
// =====================================================================================================
// Step 1 on every rank.  d_iq points at the shard's first own sample; when has_halo != 0 the sample that precedes
// the shard in the capture is stored immediately before it (d_iq[-1]).  Keeps the tile table in the arena for step 2.
// h_summary = {last_cls, last_len, whole, init_cls}: the shard's closing run and (rank 0) the digitizer's initial state.
extern "C" int urh_shard_dense(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, int has_halo, float noise_mag, int mod_type,
                               float center, uint16_t tolerance, uint8_t bits_per_symbol, float center_spacing,
                               float* d_qad_out, int64_t* h_summary) {
    if (n <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "empty shard");
    if (mod_type != URH_MOD_ASK && mod_type != URH_MOD_FSK) URH_FAIL(ctx, URH_ERR_INVALID, "sharded path: ASK / FSK only (PSK is a serial recurrence)");
    if (urh_iq_bytes(dtype) == 0) URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");
    urh_arena_reset(ctx);
    UrhClassify cls;
    URH_CHECK(fill_classify(ctx, &cls, mod_type, center, bits_per_symbol, center_spacing));
    const UrhDemodParams dp = make_demod_params(noise_mag, mod_type, dtype);
    const int tol = tolerance;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = stage_cap_for(tol);
    UrhTileSummary* tiles;
    uint32_t* staging;
    int16_t* d_init;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    URH_CHECK(urh_arena(ctx, 8, &d_init));
    URH_CUDA(ctx, cudaMemsetAsync(d_init, 0, 16, ctx->stream));
    const int c0 = host_classify(0.0f, cls);
    if (mod_type == URH_MOD_ASK)
        URH_CHECK((launch_dense_iq_m<URH_MOD_ASK, true>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, tol, tiles, staging, cap, d_init, c0, has_halo)));
    else
        URH_CHECK((launch_dense_iq_m<URH_MOD_FSK, true>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, tol, tiles, staging, cap, d_init, c0, has_halo)));
    ctx->shard_tiles = tiles;
    ctx->shard_staging = staging;
    ctx->shard_cap = cap;
    ctx->shard_n = n;
    ctx->shard_tol = tol;
    URH_CHECK(urh_shard_run_total(ctx, n, tiles, h_summary));
    int16_t init16 = 0;
    URH_CUDA(ctx, cudaMemcpyAsync(&init16, d_init, sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_summary[3] = init16;
    return URH_OK;
}

// Step 1 for a shard that is already demodulated (the center became known only after the demodulation pass): the
// digitizer's dense pass over qad.  Classification is per sample, so no halo is involved; same h_summary as urh_shard_dense.
extern "C" int urh_shard_dense_qad(urh_ctx* ctx, const float* d_qad, int64_t n, int mod_type, float center, uint16_t tolerance,
                                   uint8_t bits_per_symbol, float center_spacing, int64_t* h_summary) {
    if (n <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "empty shard");
    urh_arena_reset(ctx);
    UrhClassify cls;
    URH_CHECK(fill_classify(ctx, &cls, mod_type, center, bits_per_symbol, center_spacing));
    const int tol = tolerance;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = stage_cap_for(tol);
    UrhTileSummary* tiles;
    uint32_t* staging;
    int16_t* d_init;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    URH_CHECK(urh_arena(ctx, 8, &d_init));
    URH_CUDA(ctx, cudaMemsetAsync(d_init, 0, 16, ctx->stream));
    const int vec_in = (((uintptr_t)d_qad % 8) == 0) ? 1 : 0;
    URH_PROF_BEGIN(ctx);
    const unsigned grid = (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK);
    if (cls.order == 2)
        URH_LAUNCH(ctx, (k_dense_f32<SrcQad2, float>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_qad, n, vec_in, cls, tol, tiles, staging, cap,
                   d_init, host_classify(0.0f, cls));
    else
        URH_LAUNCH(ctx, (k_dense_f32<SrcQad, float>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_qad, n, vec_in, cls, tol, tiles, staging, cap,
                   d_init, host_classify(0.0f, cls));
    URH_PROF_END(ctx);
    ctx->shard_tiles = tiles;
    ctx->shard_staging = staging;
    ctx->shard_cap = cap;
    ctx->shard_n = n;
    ctx->shard_tol = tol;
    URH_CHECK(urh_shard_run_total(ctx, n, tiles, h_summary));
    int16_t init16 = 0;
    URH_CUDA(ctx, cudaMemcpyAsync(&init16, d_init, sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_summary[3] = init16;
    return URH_OK;
}

// ---- one-call paths: every stage enqueued on the context stream, ONE synchronisation at the end -----------------------------
struct CenterPlan;
int urh_center_chain(urh_ctx* ctx, const float* d_qad, int64_t n, const UrhTileStats* ts, int64_t max_size, int rank, int world,
                     CenterPlan** d_plan_out);   // center.cu
int urh_center_plan_result(urh_ctx* ctx, const CenterPlan* plan, const float** d_centerf, const double** d_center, const int** d_state);

// Sharded demod + digitize for a KNOWN center: dense pass over this rank's shard, then the tile-level finish with
// its three 16-byte exchanges on the stream (finish.cu).  d_qad_in != NULL: the shard is already demodulated, digitize from it.
// Every rank ends with the rows of its own shard (urh_fetch_pulses).
extern "C" int urh_shard_digitize(urh_ctx* ctx, const void* d_iq, int dtype, const float* d_qad_in, int64_t n, int has_halo,
                                  float noise_mag, int mod_type, float center, uint16_t tolerance, uint32_t samples_per_symbol,
                                  uint8_t bits_per_symbol, float center_spacing, float* d_qad_out, int64_t global_offset,
                                  int64_t n_total, int64_t* k) {
    if (!k) return URH_ERR_INVALID;
    *k = 0;
    ctx->pulses_k = 0;
    if (n <= 0) URH_FAIL(ctx, URH_ERR_INVALID, "empty shard");
    if (mod_type != URH_MOD_ASK && mod_type != URH_MOD_FSK) URH_FAIL(ctx, URH_ERR_INVALID, "sharded path: ASK / FSK only");
    if (!d_qad_in && urh_iq_bytes(dtype) == 0) URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");
    if (!ctx->nccl_comm) URH_FAIL(ctx, URH_ERR_INVALID, "NCCL communicator not initialised (urh_nccl_init)");
    urh_arena_reset(ctx);
    UrhClassify cls;
    URH_CHECK(fill_classify(ctx, &cls, mod_type, center, bits_per_symbol, center_spacing));
    const int tol = tolerance;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = stage_cap_for(tol);
    UrhTileSummary* tiles;
    uint32_t* staging;
    int16_t* d_init;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    URH_CHECK(urh_arena(ctx, 8, &d_init));
    URH_CUDA(ctx, cudaMemsetAsync(d_init, 0, 16, ctx->stream));
    const int c0 = host_classify(0.0f, cls);
    if (d_qad_in) {
        const int vec_in = (((uintptr_t)d_qad_in % 8) == 0) ? 1 : 0;
        const unsigned grid = (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK);
        URH_PROF_BEGIN(ctx);
        if (cls.order == 2)
            URH_LAUNCH(ctx, (k_dense_f32<SrcQad2, float>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_qad_in, n, vec_in, cls, tol, tiles, staging, cap,
                       d_init, c0, (const float*)nullptr);
        else
            URH_LAUNCH(ctx, (k_dense_f32<SrcQad, float>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_qad_in, n, vec_in, cls, tol, tiles, staging, cap,
                       d_init, c0, (const float*)nullptr);
        URH_PROF_END(ctx);
    } else {
        const UrhDemodParams dp = make_demod_params(noise_mag, mod_type, dtype);
        if (mod_type == URH_MOD_ASK)
            URH_CHECK((launch_dense_iq_m<URH_MOD_ASK, true>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, tol, tiles, staging, cap, d_init, c0, has_halo)));
        else
            URH_CHECK((launch_dense_iq_m<URH_MOD_FSK, true>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, tol, tiles, staging, cap, d_init, c0, has_halo)));
    }
    return urh_finish_shard(ctx, n, tol, mod_type == URH_MOD_ASK, samples_per_symbol, tiles, staging, cap, d_init, global_offset, n_total, k);
}

// demod (ASK / FSK) + capture-wide detect_center + digitize (binary symbols) in one call (BASELINE configs[1]).
// sharded != 0: this rank's shard of a capture spread over the context's NCCL communicator (has_halo as urh_shard_dense).
// *center_state: 0 = detect_center finds no center (None; *k = 0), 1 = *center is valid, 2 = the device could not decide
// (a tie between histogram peaks whose order numpy's argsort defines, or more than 6000 bins): d_qad_out is valid, the
// caller finishes through the stepwise entry points (urh_center_window_stats / urh_center_histogram_tiles / urh_grab_pulse_lens).
static int demod_center_digitize_impl(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, int has_halo, float noise_mag, int mod_type,
                                      uint16_t tolerance, uint32_t samples_per_symbol, int64_t max_size, float* d_qad_out, bool sharded,
                                      int64_t global_offset, int64_t n_total, double* center, int* center_state, int64_t* k,
                                      const void* h_iq = nullptr, int64_t chunk_samples = 0) {
    if (!k || !center || !center_state) return URH_ERR_INVALID;
    *k = 0;
    *center = 0.0;
    *center_state = 0;
    ctx->pulses_k = 0;
    if (n <= 2 || (mod_type != URH_MOD_ASK && mod_type != URH_MOD_FSK) || !d_qad_out)
        URH_FAIL(ctx, URH_ERR_INVALID, "demod_center_digitize: ASK/FSK, n > 2 and a qad buffer are required");
    if (urh_iq_bytes(dtype) == 0) URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");
    if (sharded && !ctx->nccl_comm) URH_FAIL(ctx, URH_ERR_INVALID, "NCCL communicator not initialised (urh_nccl_init)");
    urh_arena_reset(ctx);
    URH_TL_RESET(ctx);
    URH_TL_MARK(ctx, "step start");
    const UrhDemodParams dp = make_demod_params(noise_mag, mod_type, dtype);
    UrhClassify cls;
    memset(&cls, 0, sizeof(cls));
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    UrhTileStats* ts;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &ts));
    // h_iq != NULL: the capture is in (pinned) host memory.  It is uploaded in chunks on the copy stream and every chunk is
    // demodulated as soon as it has landed, so the demodulation pass hides behind the PCIe transfer.
    const int64_t chunk_tiles = (h_iq && chunk_samples > 0) ? (chunk_samples >= URH_TILE ? chunk_samples / URH_TILE : 1) : ntiles;
    const size_t sample_bytes = (size_t)urh_iq_bytes(dtype);
    int chunk_no = 0;
    for (int64_t t0 = 0; t0 < ntiles; t0 += chunk_tiles, chunk_no++) {
        const int64_t t1 = (t0 + chunk_tiles < ntiles) ? t0 + chunk_tiles : ntiles;
        if (h_iq) {
            const int64_t s0 = t0 * URH_TILE, s1 = (t1 * URH_TILE < n) ? t1 * URH_TILE : n;
            cudaEvent_t ev = ctx->ev_copy[chunk_no & 1];
            URH_CUDA(ctx, cudaMemcpyAsync((char*)d_iq + (size_t)s0 * sample_bytes, (const char*)h_iq + (size_t)s0 * sample_bytes,
                                          (size_t)(s1 - s0) * sample_bytes, cudaMemcpyHostToDevice, ctx->copy_stream[0]));
            URH_CUDA(ctx, cudaEventRecord(ev, ctx->copy_stream[0]));
            URH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev, 0));
        }
        if (mod_type == URH_MOD_ASK)
            URH_CHECK((launch_dense_iq_m<URH_MOD_ASK, false>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, 0, nullptr, nullptr, 0, nullptr, 0, has_halo, ts, t0, t1)));
        else
            URH_CHECK((launch_dense_iq_m<URH_MOD_FSK, false>(ctx, dtype, d_iq, n, dp, d_qad_out, cls, 0, nullptr, nullptr, 0, nullptr, 0, has_halo, ts, t0, t1)));
    }
    CenterPlan* plan = nullptr;
    URH_CHECK(urh_center_chain(ctx, d_qad_out, n, ts, max_size, sharded ? ctx->nccl_rank : 0, sharded ? ctx->nccl_world : 1, &plan));
    const float* d_centerf;
    const double* d_center;
    const int* d_state;
    urh_center_plan_result(ctx, plan, &d_centerf, &d_center, &d_state);
    // digitizer pass over qad, threshold read from device memory
    cls.noise_value = urh_noise_value(mod_type);
    cls.order = 2;
    const int tol = tolerance;
    const int cap = stage_cap_for(tol);
    UrhTileSummary* tiles;
    uint32_t* staging;
    int16_t* d_init;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    URH_CHECK(urh_arena(ctx, 8, &d_init));
    URH_CUDA(ctx, cudaMemsetAsync(d_init, 0, 16, ctx->stream));
    const int vec_in = (((uintptr_t)d_qad_out % 8) == 0) ? 1 : 0;
    URH_LAUNCH(ctx, (k_dense_f32<SrcQad2, float>), (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK), URH_WARPS_PER_BLOCK * 32, 0,
               (const float*)d_qad_out, n, vec_in, cls, tol, tiles, staging, cap, d_init, 0, d_centerf, (const UrhTileStats*)ts);
    URH_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 40, d_center, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 41, d_state, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    int64_t rows = 0;
    if (sharded)
        URH_CHECK(urh_finish_shard(ctx, n, tol, mod_type == URH_MOD_ASK, samples_per_symbol, tiles, staging, cap, d_init, global_offset, n_total, &rows));
    else
        URH_CHECK(urh_finish_local(ctx, n, tol, mod_type == URH_MOD_ASK, samples_per_symbol, tiles, staging, cap, d_init, &rows));
    // the finish synchronised the stream: the two scalars have landed
    memcpy(center, ctx->h_mail + 40, sizeof(double));
    int st = 0;
    memcpy(&st, ctx->h_mail + 41, sizeof(int));
    *center_state = st;
    if (st != 1) {
        ctx->pulses_k = 0;
        rows = 0;
    }
    *k = rows;
    return URH_OK;
}

extern "C" int urh_demod_center_digitize(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, float noise_mag, int mod_type,
                                         uint16_t tolerance, uint32_t samples_per_symbol, int64_t max_size, float* d_qad_out,
                                         double* center, int* center_state, int64_t* k) {
    return demod_center_digitize_impl(ctx, d_iq, dtype, n, 0, noise_mag, mod_type, tolerance, samples_per_symbol, max_size, d_qad_out, false,
                                      0, n, center, center_state, k);
}

// The same step fed from HOST memory (pinned for a truly asynchronous copy): the IQ samples are uploaded into d_iq_scratch in
// chunks of `chunk_samples` on the copy stream while the compute stream demodulates the chunks that have landed (streaming
// ingest).  chunk_samples <= 0: 2^24.  Works for every IQArray dtype (int8 / int16 captures move 4x / 2x fewer bytes).
extern "C" int urh_demod_center_digitize_host(urh_ctx* ctx, const void* h_iq, int dtype, int64_t n, float noise_mag, int mod_type,
                                              uint16_t tolerance, uint32_t samples_per_symbol, int64_t max_size, int64_t chunk_samples,
                                              void* d_iq_scratch, float* d_qad_out, double* center, int* center_state, int64_t* k) {
    if (!h_iq || !d_iq_scratch) return URH_ERR_INVALID;
    // the copy stream must not run ahead of work already queued on the compute stream that still reads the scratch buffer
    URH_CUDA(ctx, cudaEventRecord(ctx->ev_comp[0], ctx->stream));
    URH_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream[0], ctx->ev_comp[0], 0));
    return demod_center_digitize_impl(ctx, d_iq_scratch, dtype, n, 0, noise_mag, mod_type, tolerance, samples_per_symbol, max_size, d_qad_out,
                                      false, 0, n, center, center_state, k, h_iq, chunk_samples > 0 ? chunk_samples : ((int64_t)1 << 24));
}

// ... and for one shard of a sharded capture (the halo sample, if any, must already sit at d_iq_scratch[-1])
extern "C" int urh_shard_demod_center_digitize_host(urh_ctx* ctx, const void* h_iq, int dtype, int64_t n, int has_halo, float noise_mag,
                                                    int mod_type, uint16_t tolerance, uint32_t samples_per_symbol, int64_t max_size,
                                                    int64_t chunk_samples, void* d_iq_scratch, float* d_qad_out, int64_t global_offset,
                                                    int64_t n_total, double* center, int* center_state, int64_t* k) {
    if (!h_iq || !d_iq_scratch) return URH_ERR_INVALID;
    URH_CUDA(ctx, cudaEventRecord(ctx->ev_comp[0], ctx->stream));
    URH_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream[0], ctx->ev_comp[0], 0));
    return demod_center_digitize_impl(ctx, d_iq_scratch, dtype, n, has_halo, noise_mag, mod_type, tolerance, samples_per_symbol, max_size,
                                      d_qad_out, true, global_offset, n_total, center, center_state, k, h_iq,
                                      chunk_samples > 0 ? chunk_samples : ((int64_t)1 << 24));
}

extern "C" int urh_shard_demod_center_digitize(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, int has_halo, float noise_mag,
                                               int mod_type, uint16_t tolerance, uint32_t samples_per_symbol, int64_t max_size,
                                               float* d_qad_out, int64_t global_offset, int64_t n_total, double* center,
                                               int* center_state, int64_t* k) {
    return demod_center_digitize_impl(ctx, d_iq, dtype, n, has_halo, noise_mag, mod_type, tolerance, samples_per_symbol, max_size, d_qad_out,
                                      true, global_offset, n_total, center, center_state, k);
}

struct UrhShardState {
    UrhCandidates cand;
    UrhFireState fs;
    int16_t* d_prev;
};
static UrhShardState* shard_state(urh_ctx* ctx) {
    if (!ctx->shard_state) ctx->shard_state = calloc(1, sizeof(UrhShardState));
    return (UrhShardState*)ctx->shard_state;
}

// Step 2 on every rank, after the summaries were exchanged: carry_* describe the run that ends right before this
// shard (fold of the preceding shards' summaries; carry_valid = 0 on the first shard).  Positions are global.
// *last_cand_cls = class of the shard's last candidate (meaningful when *count > 0).
extern "C" int urh_shard_candidates(urh_ctx* ctx, int carry_valid, int carry_cls, int64_t carry_len, int64_t global_offset,
                                    int64_t* count, const int64_t** d_pos, const int16_t** d_cls, int* last_cand_cls) {
    if (!ctx->shard_tiles) URH_FAIL(ctx, URH_ERR_INVALID, "urh_shard_dense must precede urh_shard_candidates");
    UrhShardCarry in;
    in.valid = carry_valid; in.cls = carry_cls; in.len = carry_len;
    UrhShardState* S = shard_state(ctx);
    URH_CHECK(urh_collect_candidates_shard(ctx, ctx->shard_n, ctx->shard_tol, (const UrhTileSummary*)ctx->shard_tiles,
                                           (const uint32_t*)ctx->shard_staging, ctx->shard_cap, in, global_offset, &S->cand));
    *count = S->cand.count;
    if (d_pos) *d_pos = S->cand.pos;
    if (d_cls) *d_cls = S->cand.cls;
    if (last_cand_cls) {
        *last_cand_cls = 0;
        if (S->cand.count > 0) {
            int16_t v = 0;
            URH_CUDA(ctx, cudaMemcpyAsync(&v, S->cand.cls + S->cand.count - 1, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
            URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            *last_cand_cls = v;
        }
    }
    ctx->shard_tiles = nullptr;
    return URH_OK;
}

// host copies of the candidate table urh_shard_candidates left on the device
extern "C" int urh_fetch_candidates(urh_ctx* ctx, int64_t* h_pos, int16_t* h_cls, int64_t count) {
    UrhShardState* S = shard_state(ctx);
    if (count < 0 || count > S->cand.count) URH_FAIL(ctx, URH_ERR_INVALID, "fetch_candidates: count exceeds the table");
    if (count == 0) return URH_OK;
    URH_CUDA(ctx, cudaMemcpyAsync(h_pos, S->cand.pos, (size_t)count * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaMemcpyAsync(h_cls, S->cand.cls, (size_t)count * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return URH_OK;
}

// Step 3 (distributed finish): prev_cls = class of the last candidate of the preceding shards (the digitizer's
// initial state on the first shard).  Returns the number of firings and the position of the last one (-1: none).
extern "C" int urh_shard_fire(urh_ctx* ctx, int prev_cls, int64_t* fired, int64_t* last_fired_pos) {
    UrhShardState* S = shard_state(ctx);
    URH_CHECK(urh_arena(ctx, 8, &S->d_prev));
    const int16_t v = (int16_t)prev_cls;
    URH_CUDA(ctx, cudaMemcpyAsync(S->d_prev, &v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    URH_CHECK(urh_fire_stage(ctx, S->cand, S->d_prev, &S->fs, last_fired_pos));
    *fired = S->fs.F;
    return URH_OK;
}

// Step 4: this shard's rows (merged locally; equal states across a shard edge are merged by the consumer).
// prev_fired_pos = position of the last firing in the preceding shards (-1: none); emit_tail on the last shard only.
extern "C" int urh_shard_rows(urh_ctx* ctx, int64_t n_total, uint16_t tolerance, int mod_type, uint32_t samples_per_symbol,
                              int64_t prev_fired_pos, int emit_tail, int64_t* k) {
    UrhShardState* S = shard_state(ctx);
    return urh_rows_stage(ctx, S->fs, n_total, tolerance, mod_type == URH_MOD_ASK, samples_per_symbol, prev_fired_pos, emit_tail != 0, k);
}

// Step 3 on the gathering rank: the concatenated candidate tables of all shards -> pulse table of the whole capture.
extern "C" int urh_pulses_from_table(urh_ctx* ctx, const int64_t* d_pos, const int16_t* d_cls, int64_t count, int64_t n_total,
                                     uint16_t tolerance, int mod_type, uint32_t samples_per_symbol, int init_cls, int64_t* k) {
    urh_arena_reset(ctx);
    int16_t* d_init;
    URH_CHECK(urh_arena(ctx, 8, &d_init));
    const int16_t v = (int16_t)init_cls;
    URH_CUDA(ctx, cudaMemcpyAsync(d_init, &v, sizeof(v), cudaMemcpyHostToDevice, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    UrhCandidates cand;
    cand.count = count;
    cand.pos = (int64_t*)d_pos;
    cand.cls = (int16_t*)d_cls;
    cand.last_cls = 0;
    cand.last_len = 0;
    return urh_pulses_from_candidates(ctx, n_total, tolerance, mod_type == URH_MOD_ASK, samples_per_symbol, cand, d_init, k);
}

extern "C" int urh_fetch_pulses(urh_ctx* ctx, int64_t* h_rows, int64_t k) {
    if (k < 0 || k > ctx->pulses_k) URH_FAIL(ctx, URH_ERR_INVALID, "fetch_pulses: k=%lld exceeds last result %lld", (long long)k, (long long)ctx->pulses_k);
    if (k == 0) return URH_OK;
    URH_CUDA(ctx, cudaMemcpyAsync(h_rows, ctx->pulses, (size_t)k * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return URH_OK;
}

extern "C" int urh_pulses_device_ptr(urh_ctx* ctx, const int64_t** d_rows, int64_t* k) {
    if (d_rows) *d_rows = ctx->pulses;
    if (k) *k = ctx->pulses_k;
    return URH_OK;
}


// ---- diagnostic: packed-division self test (tests/test_gpu_packed_div.py) ---------------------------
__global__ void k_selftest_div(uint64_t seed, int64_t count, unsigned long long* mismatches, unsigned long long* tested) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0, ok = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint64_t h = seed + (uint64_t)i * 0x9E3779B97F4A7C15ull;
        h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
        h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
        h ^= h >> 31;
        uint64_t g = h * 0xD6E8FEB86659FD93ull + 0x632BE59BD9B4E019ull;
        g ^= g >> 29;
        // exponents uniformly inside the window, random mantissas; a few structured cases
        const uint32_t span = URH_DIVWIN_HI - URH_DIVWIN_LO;
        uint32_t ea = URH_DIVWIN_LO + (uint32_t)((h >> 40) % span), eb = URH_DIVWIN_LO + (uint32_t)((g >> 40) % span);
        uint32_t ma = (uint32_t)h & 0x7fffff, mb = (uint32_t)g & 0x7fffff;
        if ((i & 15) == 1) ma = 0;
        if ((i & 15) == 2) mb = 0;
        if ((i & 15) == 3) { ma = 0x7fffff; }
        if ((i & 15) == 4) { mb = 0x7fffff; }
        if ((i & 15) == 5) { eb = ea; }
        float a0 = __uint_as_float((ea << 23) | ma), b0 = __uint_as_float((eb << 23) | mb);
        float a1 = __uint_as_float(((URH_DIVWIN_LO + (uint32_t)((g >> 12) % span)) << 23) | ((uint32_t)(h >> 9) & 0x7fffff));
        float b1 = __uint_as_float(((URH_DIVWIN_LO + (uint32_t)((h >> 12) % span)) << 23) | ((uint32_t)(g >> 9) & 0x7fffff));
        if ((i & 15) == 6) { a1 = __uint_as_float((URH_DIVWIN_LO << 23)); b1 = __uint_as_float(((URH_DIVWIN_HI - 1) << 23) | 0x7fffff); }
        if ((i & 15) == 8) { b1 = __uint_as_float((URH_DIVWIN_LO << 23)); a1 = __uint_as_float(((URH_DIVWIN_HI - 1) << 23) | 0x7fffff); }
        if ((i & 63) == 7) a1 = 0.0f;
        const float2 q = urh_div2_window(make_float2(a0, a1), make_float2(b0, b1));
        const float r0 = __fdiv_rn(a0, b0), r1 = __fdiv_rn(a1, b1);
        bad += (__float_as_uint(q.x) != __float_as_uint(r0)) + (__float_as_uint(q.y) != __float_as_uint(r1));
        ok += 2;
    }
    atomicAdd(mismatches, bad);
    atomicAdd(tested, ok);
}

extern "C" int urh_selftest_packed_div(urh_ctx* ctx, uint64_t seed, int64_t count, int64_t* mismatches, int64_t* tested) {
    urh_arena_reset(ctx);
    unsigned long long* d;
    URH_CHECK(urh_arena(ctx, 2, &d));
    URH_CUDA(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
    URH_LAUNCH(ctx, k_selftest_div, (unsigned)(ctx->sm_count * 8), 256, 0, seed, count, d, d + 1);
    int64_t h[2];
    URH_CHECK(urh_read_i64(ctx, (const int64_t*)d, 2, h));
    *mismatches = h[0];
    *tested = h[1];
    return URH_OK;
}
