// detect_center on the GPU (AutoInterpretation.detect_center, AutoInterpretation.py:226-277).
//
//   rect = x[x > -4]; rect = rect[int(0.05*len) : int(0.95*len)] (optionally [:max_size])   -- trimming by RANK among kept samples
//   bins = np.arange(min(rect), max(rect) + var(rect), var(rect)); y = np.histogram(rect, bins)   -> peak picking (host)
//
// Everything is organised around the dense pass's tiles (URH_TILE samples, one warp each):
//   1. a table of per-tile {count, min, max, sum, sumsq} of the kept samples (UrhTileStats).  The FSK/ASK demodulator
//      produces it for free in the pass that writes qad (urh_afp_demod_tiles, digitize.cu); for an array that is already
//      demodulated k_tile_stats_f32 reads it once.
//   2. rank prefix over the tile counts -> the two tiles the rank window cuts; window statistics = table entries of the
//      interior tiles + a rank-exact re-read of the (at most two) cut tiles.  No pass over the samples.
//   3. histogram: ONE pass over qad.  Bin edges become float thresholds (exact, see k_hist_edges), every thread counts
//      its currently popular bins in registers and only misses touch the shared-memory histogram.
#include "dense.cuh"
#include "scan.cuh"

#include <math.h>

struct CenStats {
    double sum, sumsq;
    float mn, mx;
    long long cnt;
};

// ---- 1. tile table from an already-demodulated array ------------------------------------------------------------------
__global__ void __launch_bounds__(URH_WARPS_PER_BLOCK * 32)
k_tile_stats_f32(const float* __restrict__ x, int64_t n, int64_t ntiles, UrhTileStats* __restrict__ ts) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * URH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (tile >= ntiles) return;
    const int64_t base = tile * URH_TILE;
    UrhStatAcc acc;
    acc.init();
    if (base + URH_TILE <= n && (((uintptr_t)x) & 15) == 0) {
        const float4* p = (const float4*)(x + base) + lane;
        constexpr int ITERS = URH_TILE / 128;
        float4 cur[4], nxt[4];
#pragma unroll
        for (int j = 0; j < 4; j++) cur[j] = __ldg(p + j * 32);
        for (int it = 0; it < ITERS; it += 4) {
            if (it + 4 < ITERS) {
#pragma unroll
                for (int j = 0; j < 4; j++) nxt[j] = __ldg(p + (it + 4 + j) * 32);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                acc.add(cur[j].x); acc.add(cur[j].y); acc.add(cur[j].z); acc.add(cur[j].w);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) cur[j] = nxt[j];
        }
    } else {
        for (int j = lane; j < URH_TILE; j += 32)
            if (base + j < n) acc.add(x[base + j]);
    }
    acc.store(ts + tile, lane);
}

// ---- 2. rank prefix, window tiles, window statistics --------------------------------------------------------------------
__global__ void k_tile_counts(const UrhTileStats* __restrict__ ts, int64_t ntiles, int64_t* __restrict__ prefix) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles) prefix[t] = ts[t].cnt;
}

// Tile t holds ranks [prefix[t], prefix[t+1]).  win[0] = tile of rank r0, win[1] = tile of rank r1-1 (r0 < r1 <= total);
// win[2], win[3] = those tiles again if the window cuts them (covers them only partly), else -1.  Pre-set to -1.
__global__ void k_find_window_tiles(const int64_t* __restrict__ prefix, int64_t ntiles, int64_t r0, int64_t r1, int64_t* __restrict__ win) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int64_t a = prefix[t], b = prefix[t + 1];
    if (b <= a) return;
    const bool covered = a >= r0 && b <= r1;
    const bool has_r0 = a <= r0 && r0 < b, has_r1 = a <= r1 - 1 && r1 - 1 < b;
    if (has_r0) { win[0] = t; win[2] = covered ? -1 : t; }
    if (has_r1) { win[1] = t; win[3] = (covered || has_r0) ? -1 : t; }
}

// block-wide exclusive rank of each thread's first kept sample inside one tile (8 consecutive samples per thread)
#define CEN_PER (URH_TILE / 256)
__device__ __forceinline__ int64_t cen_tile_ranks(const float* __restrict__ x, int64_t n, int64_t t, int64_t tile_rank0, float (&v)[CEN_PER],
                                                  int* s_pre) {
    const int64_t base = t * URH_TILE + (int64_t)threadIdx.x * CEN_PER;
    int mine = 0;
#pragma unroll
    for (int j = 0; j < CEN_PER; j++) {
        v[j] = (base + j < n) ? x[base + j] : -5.0f;
        mine += (v[j] > -4.0f) ? 1 : 0;
    }
    s_pre[threadIdx.x] = mine;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        int add = 0;
        if (threadIdx.x >= off) add = s_pre[threadIdx.x - off];
        __syncthreads();
        s_pre[threadIdx.x] += add;
        __syncthreads();
    }
    return tile_rank0 + s_pre[threadIdx.x] - mine;
}

__device__ __forceinline__ void cen_block_fold(double sum, double sq, float mn, float mx, long long cnt, CenStats* out) {
    __shared__ double s_sum[256], s_sq[256];
    __shared__ float s_mn[256], s_mx[256];
    __shared__ long long s_cnt[256];
    s_sum[threadIdx.x] = sum; s_sq[threadIdx.x] = sq; s_mn[threadIdx.x] = mn; s_mx[threadIdx.x] = mx; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
            s_sq[threadIdx.x] += s_sq[threadIdx.x + off];
            s_mn[threadIdx.x] = fminf(s_mn[threadIdx.x], s_mn[threadIdx.x + off]);
            s_mx[threadIdx.x] = fmaxf(s_mx[threadIdx.x], s_mx[threadIdx.x + off]);
            s_cnt[threadIdx.x] += s_cnt[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        CenStats r;
        r.sum = s_sum[0]; r.sumsq = s_sq[0]; r.mn = s_mn[0]; r.mx = s_mx[0]; r.cnt = s_cnt[0];
        *out = r;
    }
}

// rank-exact partial statistics of the cut tiles win[2], win[3]; one block each
__global__ void __launch_bounds__(256) k_cut_tile_stats(const float* __restrict__ x, int64_t n, const int64_t* __restrict__ prefix,
                                                       const int64_t* __restrict__ win, int64_t r0, int64_t r1, CenStats* __restrict__ out) {
    __shared__ int s_pre[256];
    const int64_t t = win[2 + blockIdx.x];
    double sum = 0.0, sq = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    if (t >= 0) {
        float v[CEN_PER];
        int64_t rank = cen_tile_ranks(x, n, t, prefix[t], v, s_pre);
#pragma unroll
        for (int j = 0; j < CEN_PER; j++) {
            if (v[j] > -4.0f) {
                if (rank >= r0 && rank < r1) {
                    sum += (double)v[j];
                    sq += (double)v[j] * (double)v[j];
                    mn = fminf(mn, v[j]);
                    mx = fmaxf(mx, v[j]);
                    cnt++;
                }
                rank++;
            }
        }
    }
    cen_block_fold(sum, sq, mn, mx, cnt, out + blockIdx.x);
}

// tiles entirely inside the rank window: fold the table; grid-stride, one partial per block
__global__ void __launch_bounds__(256) k_interior_tile_stats(const UrhTileStats* __restrict__ ts, const int64_t* __restrict__ prefix,
                                                            int64_t ntiles, int64_t r0, int64_t r1, CenStats* __restrict__ partial) {
    double sum = 0.0, sq = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < ntiles; t += (int64_t)gridDim.x * 256) {
        const int64_t a = prefix[t], b = prefix[t + 1];
        if (b > a && a >= r0 && b <= r1) {
            const UrhTileStats v = ts[t];
            sum += v.sum; sq += v.sumsq; mn = fminf(mn, v.mn); mx = fmaxf(mx, v.mx); cnt += v.cnt;
        }
    }
    cen_block_fold(sum, sq, mn, mx, cnt, partial + blockIdx.x);
}

__global__ void __launch_bounds__(256) k_center_fold(const CenStats* __restrict__ partial, int64_t count, CenStats* __restrict__ out) {
    double sum = 0.0, sq = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    for (int64_t t = threadIdx.x; t < count; t += 256) {
        const CenStats p = partial[t];
        sum += p.sum; sq += p.sumsq; mn = fminf(mn, p.mn); mx = fmaxf(mx, p.mx); cnt += p.cnt;
    }
    cen_block_fold(sum, sq, mn, mx, cnt, out);
}

// Rank prefix over a tile table; leaves {x, ts, prefix, n} in ctx for the window / histogram calls.
int urh_center_tiles_begin(urh_ctx* ctx, const float* d_x, const UrhTileStats* ts, int64_t n, int64_t* h_total) {
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    int64_t *prefix, *d_total;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles + 1, &prefix));
    URH_CHECK(urh_arena(ctx, 4, &d_total));
    URH_LAUNCH(ctx, k_tile_counts, (unsigned)urh_div_up(ntiles, 256), 256, 0, ts, ntiles, prefix);
    URH_CHECK((urhscan::device_scan<int64_t, urhscan::AddI64>(ctx, prefix, ntiles, urhscan::AddI64(), (int64_t)0, true, d_total)));
    URH_CUDA(ctx, cudaMemcpyAsync(prefix + ntiles, d_total, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    URH_CHECK(urh_read_i64(ctx, d_total, 1, h_total));
    ctx->center_prefix = prefix;
    ctx->center_ts = ts;
    ctx->center_n = n;
    ctx->center_x = d_x;
    return URH_OK;
}

static int tiles_from_array(urh_ctx* ctx, const float* d_x, int64_t n, int64_t* h_total) {
    urh_arena_reset(ctx);
    ctx->center_prefix = nullptr;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    UrhTileStats* ts;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &ts));
    URH_LAUNCH(ctx, k_tile_stats_f32, (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK), URH_WARPS_PER_BLOCK * 32, 0, d_x, n, ntiles, ts);
    return urh_center_tiles_begin(ctx, d_x, ts, n, h_total);
}

static bool tiles_match(const urh_ctx* ctx, const float* d_x, int64_t n) {
    return ctx->center_prefix && ctx->center_n == n && ctx->center_x == (const void*)d_x;
}

// {count, min, max, sum, sumsq} of the kept samples whose LOCAL rank is in [r0, r1): interior tiles from the table, the
// cut tiles re-read from d_qad.  A shard passes the global window minus its rank offset (clamped to its own count).
extern "C" int urh_center_window_stats(urh_ctx* ctx, const float* d_qad, int64_t n, int64_t r0, int64_t r1, double* h_out5) {
    h_out5[0] = 0.0; h_out5[1] = INFINITY; h_out5[2] = -INFINITY; h_out5[3] = 0.0; h_out5[4] = 0.0;
    if (!tiles_match(ctx, d_qad, n)) URH_FAIL(ctx, URH_ERR_INVALID, "urh_afp_demod_tiles (same qad, same n) must precede urh_center_window_stats");
    if (r1 <= r0) return URH_OK;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int64_t* prefix = (const int64_t*)ctx->center_prefix;
    const UrhTileStats* ts = (const UrhTileStats*)ctx->center_ts;
    int64_t* d_win;
    CenStats* partial;
    CenStats* folded;
    const int nb = ctx->sm_count * 2;
    URH_CHECK(urh_arena(ctx, 4, &d_win));
    URH_CHECK(urh_arena(ctx, (size_t)nb + 4, &partial));
    URH_CHECK(urh_arena(ctx, 2, &folded));
    URH_CUDA(ctx, cudaMemsetAsync(d_win, 0xff, 4 * sizeof(int64_t), ctx->stream));
    URH_LAUNCH(ctx, k_find_window_tiles, (unsigned)urh_div_up(ntiles, 256), 256, 0, prefix, ntiles, r0, r1, d_win);
    URH_LAUNCH(ctx, k_interior_tile_stats, nb, 256, 0, ts, prefix, ntiles, r0, r1, partial);
    URH_LAUNCH(ctx, k_cut_tile_stats, 2, 256, 0, d_qad, n, prefix, d_win, r0, r1, partial + nb);
    URH_LAUNCH(ctx, k_center_fold, 1, 256, 0, partial, (int64_t)nb + 2, folded);
    CenStats st;
    URH_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, folded, sizeof(CenStats), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(&st, ctx->h_mail, sizeof(st));
    h_out5[0] = (double)st.cnt; h_out5[1] = (double)st.mn; h_out5[2] = (double)st.mx; h_out5[3] = st.sum; h_out5[4] = st.sumsq;
    return URH_OK;
}

// Stage 1 of the stand-alone detect_center: h_out = {count_valid, r0, r1, min, max, mean, var} of the rank-trimmed samples.
// Leaves the tile table in the arena for urh_center_histogram on the same array.
extern "C" int urh_center_stats(urh_ctx* ctx, const float* d_x, int64_t n, int64_t max_size, double* h_out) {
    for (int i = 0; i < 7; i++) h_out[i] = 0.0;
    if (n <= 0) return URH_OK;
    int64_t total = 0;
    URH_CHECK(tiles_from_array(ctx, d_x, n, &total));
    // rect[int(0.05 * len(rect)) : int(0.95 * len(rect))]  (Python float arithmetic, truncation)
    int64_t r0 = (int64_t)(0.05 * (double)total), r1 = (int64_t)(0.95 * (double)total);
    if (max_size >= 0 && r1 - r0 > max_size) r1 = r0 + max_size;
    h_out[0] = (double)total; h_out[1] = (double)r0; h_out[2] = (double)r1;
    if (r1 <= r0) return URH_OK;
    double w[5];
    URH_CHECK(urh_center_window_stats(ctx, d_x, n, r0, r1, w));
    if (w[0] <= 0.0) return URH_OK;
    // population variance from the double sums (np.var semantics; the reference's float32 pairwise result differs ~1e-7)
    const double mean = w[3] / w[0];
    double ss = w[4] - w[0] * mean * mean;
    if (ss < 0.0) ss = 0.0;
    h_out[3] = w[1]; h_out[4] = w[2]; h_out[5] = mean; h_out[6] = ss / w[0];
    return URH_OK;
}

// ---- 3. histogram -------------------------------------------------------------------------------------------------------
// Bin edges of np.histogram as FLOAT thresholds: a float sample f satisfies f >= edge_k (double) iff f >= ru(edge_k), the
// smallest float not below the edge, so the binning needs no double arithmetic and stays exact.
// fe[0..nbins] = ru(hmin + k*hstep); fe[nbins+1] = rd(last edge) (np.histogram closes the last bin);
// fe[nbins+2] = the smallest float that both exceeds -4 (detect_center's filter) and reaches the first edge.
__global__ void k_hist_edges(double hmin, double hstep, int64_t nbins, float* __restrict__ fe) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // edges exactly as np.arange forms them: one rounded product, one rounded sum (no FMA)
    if (k <= nbins) fe[k] = __double2float_ru(__dadd_rn(hmin, __dmul_rn((double)k, hstep)));
    if (k == nbins) {
        fe[nbins + 1] = __double2float_rd(__dadd_rn(hmin, __dmul_rn((double)nbins, hstep)));
        fe[nbins + 2] = fmaxf(__double2float_ru(hmin), nextafterf(-4.0f, 0.0f));
    }
}

// FAST: the bin guess rn((f - hmin) / hstep) in float arithmetic is off by less than half a bin (the host checks
// |edge| / hstep < 2^20), so the true bin is the guess or the one below it: one table look-up, no loop.
struct HistBins {
    const float* fe;   // thresholds (shared or global)
    float f_min, f_hi, scale, off;
    int nbins;
    __device__ __forceinline__ void load(const float* fe_, float scale_, int nbins_) {
        fe = fe_; nbins = nbins_; scale = scale_;
        f_hi = fe[nbins + 1];
        f_min = fe[nbins + 2];
        off = -fe[0] * scale;
    }
    // bin of one value, -1 when it does not count (np.histogram: half-open bins, the last one closed)
    template <bool FAST>
    __device__ __forceinline__ int bin_of(float f) const {
        const bool valid = f >= f_min && f <= f_hi;
        if (FAST) {
            const float t = fmaf(f, scale, off);
            int r = __float_as_int(t + 12582912.0f) - 0x4B400000;   // round to nearest via the 1.5 * 2^23 trick
            r = max(0, min(r, nbins));
            int k = r - ((f < fe[r]) ? 1 : 0);
            k = min(k, nbins - 1);
            return valid ? k : -1;
        }
        if (!valid) return -1;
        int k = (int)((f - fe[0]) * scale);
        k = max(0, min(k, nbins - 1));
        while (k > 0 && f < fe[k]) k--;
        while (k < nbins - 1 && f >= fe[k + 1]) k++;
        return k;
    }
};

// A demodulated capture piles its samples onto a handful of bins.  Each thread keeps four bins in registers as float
// INTERVALS [lo, hi) with a count: a sample is compared with the four intervals directly (no bin index is computed on a
// hit); only a sample that falls into none of them is binned through the table and takes over the least used entry, whose
// count goes to the histogram.
struct HistCache {
    float lo0, hi0, lo1, hi1, lo2, hi2, lo3, hi3;
    int h0, h1, h2, h3;
    unsigned c0, c1, c2, c3;
    __device__ __forceinline__ void init() {
        lo0 = lo1 = lo2 = lo3 = INFINITY;   // empty interval: nothing is >= +inf
        hi0 = hi1 = hi2 = hi3 = -INFINITY;
        h0 = h1 = h2 = h3 = 0;
        c0 = c1 = c2 = c3 = 0u;
    }
};

template <bool SMEM>
__device__ __forceinline__ void hist_bump(unsigned int* s_hist, unsigned long long* hist, int k, unsigned c) {
    if (SMEM) atomicAdd(&s_hist[k], c);
    else atomicAdd(&hist[k], (unsigned long long)c);
}

// f lies inside the histogram range but in none of the cached bins
template <bool SMEM, bool FAST>
__device__ __forceinline__ void hist_miss(HistCache& hc, const HistBins& hb, unsigned int* s_hist, unsigned long long* hist, float f) {
    const int k = hb.bin_of<FAST>(f);
    if (k < 0) return;
    // the cached interval must reproduce the validity test as well: nothing below f_min, the last bin closed at f_hi
    const float lo = fmaxf(hb.fe[k], hb.f_min);
    const float hi = (k == hb.nbins - 1) ? nextafterf(hb.f_hi, INFINITY) : hb.fe[k + 1];
    unsigned cm = hc.c0; int which = 0;
    if (hc.c1 < cm) { cm = hc.c1; which = 1; }
    if (hc.c2 < cm) { cm = hc.c2; which = 2; }
    if (hc.c3 < cm) { cm = hc.c3; which = 3; }
    const int old = which == 0 ? hc.h0 : which == 1 ? hc.h1 : which == 2 ? hc.h2 : hc.h3;
    if (cm) hist_bump<SMEM>(s_hist, hist, old, cm);
    if (which == 0) { hc.h0 = k; hc.c0 = 1u; hc.lo0 = lo; hc.hi0 = hi; }
    else if (which == 1) { hc.h1 = k; hc.c1 = 1u; hc.lo1 = lo; hc.hi1 = hi; }
    else if (which == 2) { hc.h2 = k; hc.c2 = 1u; hc.lo2 = lo; hc.hi2 = hi; }
    else { hc.h3 = k; hc.c3 = 1u; hc.lo3 = lo; hc.hi3 = hi; }
}

template <bool SMEM, bool FAST>
__device__ __forceinline__ void hist_put(HistCache& hc, const HistBins& hb, unsigned int* s_hist, unsigned long long* hist, float f) {
    // per entry: two compares chained into one predicate and a predicated increment (spelled out in PTX: the compiler
    // otherwise turns each increment into select + add + move); miss = inside the histogram's range but in no cached bin
    unsigned miss;
    asm("{\n\t.reg .pred p0, p1, p2, p3, pv;\n\t"
        "setp.ge.f32 p0, %5, %6;\n\tsetp.lt.and.f32 p0, %5, %7, p0;\n\t@p0 add.u32 %0, %0, 1;\n\t"
        "setp.ge.f32 p1, %5, %8;\n\tsetp.lt.and.f32 p1, %5, %9, p1;\n\t@p1 add.u32 %1, %1, 1;\n\t"
        "setp.ge.f32 p2, %5, %10;\n\tsetp.lt.and.f32 p2, %5, %11, p2;\n\t@p2 add.u32 %2, %2, 1;\n\t"
        "setp.ge.f32 p3, %5, %12;\n\tsetp.lt.and.f32 p3, %5, %13, p3;\n\t@p3 add.u32 %3, %3, 1;\n\t"
        "or.pred p0, p0, p1;\n\tor.pred p2, p2, p3;\n\tor.pred p0, p0, p2;\n\t"
        "setp.ge.f32 pv, %5, %14;\n\tsetp.le.and.f32 pv, %5, %15, pv;\n\t"
        "and.pred pv, pv, !p0;\n\t"
        "selp.u32 %4, 1, 0, pv;\n\t}"
        : "+r"(hc.c0), "+r"(hc.c1), "+r"(hc.c2), "+r"(hc.c3), "=r"(miss)
        : "f"(f), "f"(hc.lo0), "f"(hc.hi0), "f"(hc.lo1), "f"(hc.hi1), "f"(hc.lo2), "f"(hc.hi2), "f"(hc.lo3), "f"(hc.hi3),
          "f"(hb.f_min), "f"(hb.f_hi));
    if (miss) hist_miss<SMEM, FAST>(hc, hb, s_hist, hist, f);
}

template <bool SMEM>
__device__ __forceinline__ void hist_flush(HistCache& hc, unsigned int* s_hist, unsigned long long* hist) {
    if (hc.c0) hist_bump<SMEM>(s_hist, hist, hc.h0, hc.c0);
    if (hc.c1) hist_bump<SMEM>(s_hist, hist, hc.h1, hc.c1);
    if (hc.c2) hist_bump<SMEM>(s_hist, hist, hc.h2, hc.c2);
    if (hc.c3) hist_bump<SMEM>(s_hist, hist, hc.h3, hc.c3);
}

// Tiles strictly between win[0] and win[1] lie entirely inside the rank window: every kept sample counts, no rank
// bookkeeping, no prefix reads.  One warp per tile, grid-stride, eight 512-byte rows in flight per warp.
template <bool SMEM, bool FAST>
__global__ void __launch_bounds__(256, 4) k_hist_interior(const float* __restrict__ x, int64_t n, const int64_t* __restrict__ win,
                                                      const float* __restrict__ g_fe, float scale, int nbins,
                                                      unsigned long long* __restrict__ hist, int edges_in_smem) {
    extern __shared__ unsigned int s_dyn[];
    unsigned int* s_hist = s_dyn;                               // [nbins] when SMEM
    float* s_fe = (float*)(s_dyn + (SMEM ? nbins : 0));         // [nbins + 3] when edges_in_smem
    const int lane = threadIdx.x & 31;
    if (SMEM)
        for (int b = threadIdx.x; b < nbins; b += 256) s_hist[b] = 0u;
    if (edges_in_smem)
        for (int b = threadIdx.x; b < nbins + 3; b += 256) s_fe[b] = g_fe[b];
    __syncthreads();
    HistBins hb;
    if (FAST) hb.load(s_fe, scale, nbins);   // FAST is only launched with the table in shared memory (LDS, not generic LD)
    else hb.load(edges_in_smem ? s_fe : g_fe, scale, nbins);
    HistCache hc;
    hc.init();
    const int64_t t_first = win[0] + 1, t_end = win[1];
    const int64_t gw = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
    const bool vec = (((uintptr_t)x) & 15) == 0;
    if (win[0] >= 0) {   // < 0: empty window
        for (int64_t t = t_first + gw; t < t_end; t += nw) {
            const int64_t base = t * URH_TILE;   // interior tiles are full tiles (t < last tile)
            if (vec) {
                const float4* p = (const float4*)(x + base) + lane;
                constexpr int ITERS = URH_TILE / 128, RB = 2;   // RB rows being binned + RB rows in flight
                float4 cur[RB], nxt[RB];
#pragma unroll
                for (int j = 0; j < RB; j++) cur[j] = __ldg(p + j * 32);
#pragma unroll 1
                for (int it = 0; it < ITERS; it += RB) {
                    if (it + RB < ITERS) {
#pragma unroll
                        for (int j = 0; j < RB; j++) nxt[j] = __ldg(p + (it + RB + j) * 32);
                    }
#pragma unroll
                    for (int j = 0; j < RB; j++) {
                        hist_put<SMEM, FAST>(hc, hb, s_hist, hist, cur[j].x);
                        hist_put<SMEM, FAST>(hc, hb, s_hist, hist, cur[j].y);
                        hist_put<SMEM, FAST>(hc, hb, s_hist, hist, cur[j].z);
                        hist_put<SMEM, FAST>(hc, hb, s_hist, hist, cur[j].w);
                    }
#pragma unroll
                    for (int j = 0; j < RB; j++) cur[j] = nxt[j];
                }
            } else {
                for (int j = lane; j < URH_TILE; j += 32) hist_put<SMEM, FAST>(hc, hb, s_hist, hist, x[base + j]);
            }
        }
    }
    hist_flush<SMEM>(hc, s_hist, hist);
    if (SMEM) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += 256)
            if (s_hist[b]) atomicAdd(&hist[b], (unsigned long long)s_hist[b]);
    }
}

// the window's first and last tile (win[0], win[1]; one block each): rank-exact, straight to the global histogram
__global__ void __launch_bounds__(256) k_hist_window_ends(const float* __restrict__ x, int64_t n, const int64_t* __restrict__ prefix,
                                                         const int64_t* __restrict__ win, int64_t r0, int64_t r1,
                                                         const float* __restrict__ g_fe, float scale, int nbins,
                                                         unsigned long long* __restrict__ hist) {
    __shared__ int s_pre[256];
    const int64_t t = win[blockIdx.x];
    if (t < 0 || (blockIdx.x == 1 && t == win[0])) return;
    HistBins hb;
    hb.load(g_fe, scale, nbins);
    float v[CEN_PER];
    int64_t rank = cen_tile_ranks(x, n, t, prefix[t], v, s_pre);
#pragma unroll
    for (int j = 0; j < CEN_PER; j++) {
        const bool kept = v[j] > -4.0f;
        const int k = hb.bin_of<false>(v[j]);
        if (k >= 0 && rank >= r0 && rank < r1) atomicAdd(&hist[k], 1ull);
        rank += kept ? 1 : 0;
    }
}

// Counts for edges hmin + k*hstep, k = 0..nbins (np.arange) over the samples of LOCAL rank [r0, r1); needs the tile
// table of the same array (urh_afp_demod_tiles / urh_center_stats) in the arena.
extern "C" int urh_center_histogram_tiles(urh_ctx* ctx, const float* d_qad, int64_t n, int64_t r0, int64_t r1, double hmin,
                                          double hstep, int64_t nbins, int64_t* h_hist) {
    if (nbins <= 0) return URH_OK;
    if (!tiles_match(ctx, d_qad, n)) URH_FAIL(ctx, URH_ERR_INVALID, "urh_afp_demod_tiles (same qad, same n) must precede urh_center_histogram_tiles");
    if (nbins > (int64_t)1 << 30) URH_FAIL(ctx, URH_ERR_INVALID, "too many histogram bins");
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int64_t* prefix = (const int64_t*)ctx->center_prefix;
    unsigned long long* hist;
    float* fe;
    int64_t* d_win;
    URH_CHECK(urh_arena(ctx, (size_t)nbins, &hist));
    URH_CHECK(urh_arena(ctx, (size_t)nbins + 3, &fe));
    URH_CHECK(urh_arena(ctx, 4, &d_win));
    URH_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(unsigned long long), ctx->stream));
    URH_CUDA(ctx, cudaMemsetAsync(d_win, 0xff, 4 * sizeof(int64_t), ctx->stream));
    if (r1 > r0) {
        URH_LAUNCH(ctx, k_hist_edges, (unsigned)urh_div_up(nbins + 1, 256), 256, 0, hmin, hstep, nbins, fe);
        URH_LAUNCH(ctx, k_find_window_tiles, (unsigned)urh_div_up(ntiles, 256), 256, 0, prefix, ntiles, r0, r1, d_win);
        // shared memory (48 KB without opt-in): histogram first, then the edge table if it still fits
        const bool in_smem = nbins <= 12000;
        const int edges_smem = (in_smem && nbins <= 6000) ? 1 : 0;
        const size_t dyn = (in_smem ? (size_t)nbins * 4 : 0) + (edges_smem ? (size_t)(nbins + 3) * 4 : 0);
        const float scale = (float)(1.0 / hstep);
        // one-look-up binning needs the float guess to be good to half a bin
        const double edge_abs = fmax(fabs(hmin), fabs(hmin + (double)nbins * hstep));
        const bool fast = edges_smem && hstep > 0.0 && edge_abs / hstep < 1048576.0;
        const unsigned gs = (unsigned)min(urh_div_up(ntiles, 8), (int64_t)ctx->sm_count * 8);
        const int64_t* cw = d_win;
        const float* cfe = fe;
        if (in_smem && fast) URH_LAUNCH(ctx, (k_hist_interior<true, true>), gs, 256, dyn, d_qad, n, cw, cfe, scale, (int)nbins, hist, edges_smem);
        else if (in_smem) URH_LAUNCH(ctx, (k_hist_interior<true, false>), gs, 256, dyn, d_qad, n, cw, cfe, scale, (int)nbins, hist, edges_smem);
        else URH_LAUNCH(ctx, (k_hist_interior<false, false>), gs, 256, dyn, d_qad, n, cw, cfe, scale, (int)nbins, hist, edges_smem);
        URH_LAUNCH(ctx, k_hist_window_ends, 2, 256, 0, d_qad, n, prefix, (const int64_t*)d_win, r0, r1, (const float*)fe, scale, (int)nbins, hist);
    }
    URH_CUDA(ctx, cudaMemcpyAsync(h_hist, hist, (size_t)nbins * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return URH_OK;
}

// Stage 2 of the stand-alone detect_center.  Reuses the tile table urh_center_stats left for this array; builds it if
// the caller did something else in between.
extern "C" int urh_center_histogram(urh_ctx* ctx, const float* d_x, int64_t n, int64_t r0, int64_t r1, double hmin,
                                    double hstep, int64_t nbins, int64_t* h_hist) {
    if (nbins <= 0 || n <= 0) return URH_OK;
    if (!tiles_match(ctx, d_x, n)) {
        int64_t total = 0;
        URH_CHECK(tiles_from_array(ctx, d_x, n, &total));
    }
    return urh_center_histogram_tiles(ctx, d_x, n, r0, r1, hmin, hstep, nbins, h_hist);
}
