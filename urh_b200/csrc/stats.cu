// Auto-interpretation statistics on the GPU:
//   magnitudes / noise level  — util.get_magnitudes (util.pyx:128-136), AutoInterpretation.detect_noise_level (:60-91)
//   center detection          — AutoInterpretation.detect_center (:226-277): rank trimming, min/max/var, histogram
//   message segmentation      — auto_interpretation.segment_messages_from_magnitudes (auto_interpretation.pyx:55-111)
//   plateau lengths           — auto_interpretation.get_plateau_lengths (:179-208)
//   median filter, dB         — auto_interpretation.median_filter (:211-240), util.arr2decibel (util.pyx:38-48)
// The small, data-dependent decision logic (which chunks are quiet, which histogram bins are local maxima)
// stays on the host in urh_b200/ainterpretation/AutoInterpretation.py, exactly as in the reference; the
// sample-rate reductions run here.
#include "dense_f32.cuh"
#include "scan.cuh"
#include "sparse.cuh"

#include <math.h>

// ---- magnitudes ------------------------------------------------------------------------------------------
// float32 IQ: (double)sqrtf(fl(re*re + im*im)); integer IQ: squares and sum in (wrapping) int32, double sqrt.
template <int DT>
__device__ __forceinline__ double urh_magnitude(const void* iq, int64_t i) {
    typedef typename UrhElem<DT>::type E;
    const E* p = (const E*)iq + 2 * i;
    if (DT == URH_DT_F32) {
        const float re = (float)p[0], im = (float)p[1];
        return (double)__fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
    } else {
        const uint32_t re = (uint32_t)(int32_t)p[0], im = (uint32_t)(int32_t)p[1];
        const int32_t ssum = (int32_t)(re * re + im * im);
        return sqrt((double)ssum);
    }
}

template <int DT>
__global__ void k_magnitudes(const void* __restrict__ iq, int64_t n, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = urh_magnitude<DT>(iq, i);
}

#define URH_DISPATCH_DT(dtype, KERNEL_CALL)                         \
    switch (dtype) {                                                \
        case URH_DT_I8: { constexpr int DT = URH_DT_I8; KERNEL_CALL; } break;   \
        case URH_DT_U8: { constexpr int DT = URH_DT_U8; KERNEL_CALL; } break;   \
        case URH_DT_I16: { constexpr int DT = URH_DT_I16; KERNEL_CALL; } break; \
        case URH_DT_U16: { constexpr int DT = URH_DT_U16; KERNEL_CALL; } break; \
        case URH_DT_F32: { constexpr int DT = URH_DT_F32; KERNEL_CALL; } break; \
        default: URH_FAIL(ctx, URH_ERR_DTYPE, "Unsupported dtype");  \
    }

extern "C" int urh_get_magnitudes(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, double* d_out) {
    if (n <= 0) return URH_OK;
    const unsigned grid = (unsigned)min((int64_t)ctx->sm_count * 16, urh_div_up(n, 256));
    URH_DISPATCH_DT(dtype, URH_LAUNCH(ctx, k_magnitudes<DT>, grid, 256, 0, d_iq, n, d_out));
    return URH_OK;
}

// ---- chunk statistics for detect_noise_level -----------------------------------------------------------------
// Chunks are counted from the END of the array (AutoInterpretation.py:66-72): chunk j covers
// [n - (j+1)*cs, n - j*cs).  Each block reduces a slice of one chunk to (sum, max) in double; a second kernel
// folds the slices in a fixed order, so the result is deterministic.
#define STAT_BLOCK 256
#define STAT_SLICES 64

template <typename LOADER>
__global__ void __launch_bounds__(STAT_BLOCK) k_chunk_partial(LOADER ld, int64_t n, int64_t cs, int nchunks,
                                                              double* __restrict__ psum, double* __restrict__ pmax) {
    const int chunk = blockIdx.x / STAT_SLICES, slice = blockIdx.x % STAT_SLICES;
    const int64_t c0 = n - (int64_t)(chunk + 1) * cs;
    const int64_t per = urh_div_up(cs, STAT_SLICES);
    const int64_t s0 = c0 + (int64_t)slice * per;
    const int64_t s1 = min(s0 + per, c0 + cs);
    double sum = 0.0, mx = -1.0;
    for (int64_t i = s0 + threadIdx.x; i < s1; i += STAT_BLOCK) {
        const double m = ld(i);
        sum += m;
        mx = fmax(mx, m);   // NaN-ignoring like a sequence of `if e > maximum`
    }
    __shared__ double s_sum[STAT_BLOCK], s_max[STAT_BLOCK];
    s_sum[threadIdx.x] = sum;
    s_max[threadIdx.x] = mx;
    __syncthreads();
    for (int off = STAT_BLOCK / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
            s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        psum[blockIdx.x] = s_sum[0];
        pmax[blockIdx.x] = s_max[0];
    }
}

__global__ void k_chunk_final(const double* __restrict__ psum, const double* __restrict__ pmax, int nchunks,
                              double* __restrict__ sum, double* __restrict__ mx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    double s = 0.0, m = -1.0;
    for (int j = 0; j < STAT_SLICES; j++) {
        s += psum[c * STAT_SLICES + j];
        m = fmax(m, pmax[c * STAT_SLICES + j]);
    }
    sum[c] = s;
    mx[c] = m;
}

template <int DT>
struct LoadMagIQ {
    const void* iq;
    __device__ __forceinline__ double operator()(int64_t i) const { return urh_magnitude<DT>(iq, i); }
};
template <typename T>
struct LoadReal {
    const T* x;
    __device__ __forceinline__ double operator()(int64_t i) const { return (double)x[i]; }
};

template <typename LOADER>
static int chunk_stats(urh_ctx* ctx, LOADER ld, int64_t n, int64_t cs, int nchunks, double* h_sum, double* h_max) {
    urh_arena_reset(ctx);
    double *psum, *pmax, *sum, *mx;
    URH_CHECK(urh_arena(ctx, (size_t)nchunks * STAT_SLICES, &psum));
    URH_CHECK(urh_arena(ctx, (size_t)nchunks * STAT_SLICES, &pmax));
    URH_CHECK(urh_arena(ctx, (size_t)nchunks, &sum));
    URH_CHECK(urh_arena(ctx, (size_t)nchunks, &mx));
    URH_LAUNCH(ctx, (k_chunk_partial<LOADER>), (unsigned)(nchunks * STAT_SLICES), STAT_BLOCK, 0, ld, n, cs, nchunks, psum, pmax);
    URH_LAUNCH(ctx, k_chunk_final, (unsigned)urh_div_up(nchunks, 128), 128, 0, psum, pmax, nchunks, sum, mx);
    URH_CUDA(ctx, cudaMemcpyAsync(h_sum, sum, nchunks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaMemcpyAsync(h_max, mx, nchunks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return URH_OK;
}

// per-chunk (sum, max) of the magnitudes of an IQ capture, never materialising the float64 magnitude array
extern "C" int urh_noise_chunk_stats_iq(urh_ctx* ctx, const void* d_iq, int dtype, int64_t n, int64_t chunksize,
                                        int nchunks, double* h_sum, double* h_max) {
    if (nchunks <= 0 || chunksize <= 0 || (int64_t)nchunks * chunksize > n) URH_FAIL(ctx, URH_ERR_INVALID, "bad chunking");
    URH_DISPATCH_DT(dtype, { LoadMagIQ<DT> ld; ld.iq = d_iq; URH_CHECK(chunk_stats(ctx, ld, n, chunksize, nchunks, h_sum, h_max)); });
    return URH_OK;
}
// the same on an existing magnitude array (float32: is_f64 = 0, float64: is_f64 = 1)
extern "C" int urh_noise_chunk_stats(urh_ctx* ctx, const void* d_mags, int is_f64, int64_t n, int64_t chunksize, int nchunks,
                                     double* h_sum, double* h_max) {
    if (nchunks <= 0 || chunksize <= 0 || (int64_t)nchunks * chunksize > n) URH_FAIL(ctx, URH_ERR_INVALID, "bad chunking");
    if (is_f64) {
        LoadReal<double> ld; ld.x = (const double*)d_mags;
        return chunk_stats(ctx, ld, n, chunksize, nchunks, h_sum, h_max);
    }
    LoadReal<float> ld; ld.x = (const float*)d_mags;
    return chunk_stats(ctx, ld, n, chunksize, nchunks, h_sum, h_max);
}

// ---- detect_center: rank trimming, min / max / variance, histogram -------------------------------------------------
// rect = x[x > -4]; rect = rect[int(0.05*len) : int(0.95*len)] (optionally [:max_size])  — by RANK among the kept samples.
#define CEN_TILE 4096
__global__ void __launch_bounds__(256) k_count_valid(const float* __restrict__ x, int64_t n, int64_t* __restrict__ counts) {
    const int64_t base = (int64_t)blockIdx.x * CEN_TILE;
    int c = 0;
    for (int j = threadIdx.x; j < CEN_TILE; j += 256) {
        const int64_t i = base + j;
        if (i < n && x[i] > -4.0f) c++;
    }
    __shared__ int s[256];
    s[threadIdx.x] = c;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[blockIdx.x] = s[0];
}

struct CenStats {
    double sum, sumsq;
    float mn, mx;
    long long cnt;
};

// Each block handles one tile; within the tile the rank of an element = tile prefix + (block-local prefix).
// Elements whose rank lies in [r0, r1) contribute.  pass 0: min/max/sum (for the mean); pass 1: sum of squared
// deviations from the (double) mean; pass 2: histogram.
template <int PASS>
__global__ void __launch_bounds__(256) k_center_pass(const float* __restrict__ x, int64_t n, const int64_t* __restrict__ prefix,
                                                    int64_t r0, int64_t r1, double mean, double hmin, double hstep,
                                                    int64_t nbins, CenStats* __restrict__ partial,
                                                    unsigned long long* __restrict__ hist, int64_t ntiles, int hist_in_smem) {
    double sum = 0.0, sumsq = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    // PASS 2: block-private histogram in shared memory when the bins fit (flushed once per block)
    extern __shared__ unsigned int s_hist[];
    const bool smem_hist = (PASS == 2) && hist_in_smem;
    if (smem_hist) {
        for (int64_t b = threadIdx.x; b < nbins; b += 256) s_hist[b] = 0u;
        __syncthreads();
    }
    for (int64_t tile_idx = blockIdx.x; tile_idx < ntiles; tile_idx += gridDim.x) {
    const int64_t base = tile_idx * CEN_TILE;
    const int64_t tile_rank0 = prefix[tile_idx];
    const int64_t tile_cnt = prefix[tile_idx + 1] - tile_rank0;
    if (tile_cnt > 0 && tile_rank0 < r1 && tile_rank0 + tile_cnt > r0) {
        // thread t owns elements [t*16, t*16+16) of the tile (blocked, so ranks are monotone in t)
        const int per = CEN_TILE / 256;
        int mine = 0;
        float v[per];
#pragma unroll
        for (int j = 0; j < per; j++) {
            const int64_t i = base + threadIdx.x * per + j;
            v[j] = (i < n) ? x[i] : -5.0f;
            mine += (v[j] > -4.0f) ? 1 : 0;
        }
        __shared__ int s_pre[256];
        s_pre[threadIdx.x] = mine;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            int add = 0;
            if (threadIdx.x >= off) add = s_pre[threadIdx.x - off];
            __syncthreads();
            s_pre[threadIdx.x] += add;
            __syncthreads();
        }
        int64_t rank = tile_rank0 + s_pre[threadIdx.x] - mine;
#pragma unroll
        for (int j = 0; j < per; j++) {
            if (v[j] > -4.0f) {
                if (rank >= r0 && rank < r1) {
                    if (PASS == 0) {
                        sum += (double)v[j];
                        sumsq += (double)v[j] * (double)v[j];
                        mn = fminf(mn, v[j]);
                        mx = fmaxf(mx, v[j]);
                        cnt++;
                    } else if (PASS == 1) {
                        const double d = (double)v[j] - mean;
                        sum += d * d;
                    } else {
                        // np.histogram with explicit edges e_k = hmin + k*hstep (np.arange): right-open bins,
                        // last bin closed.  Guess the bin arithmetically, then fix against the exact edges.
                        const double a = (double)v[j];
                        int64_t k = (int64_t)floor((a - hmin) / hstep);
                        if (k < 0) k = 0;
                        if (k > nbins - 1) k = nbins - 1;
                        while (k > 0 && a < hmin + (double)k * hstep) k--;
                        while (k < nbins - 1 && a >= hmin + (double)(k + 1) * hstep) k++;
                        const double last_edge = hmin + (double)nbins * hstep;
                        if (a >= hmin && a <= last_edge) {
                            if (smem_hist) atomicAdd(&s_hist[k], 1u);
                            else atomicAdd(&hist[k], 1ull);
                        }
                    }
                }
                rank++;
            }
        }
        __syncthreads();  // s_pre is reused by the next tile
    }
    }
    if (smem_hist) {
        __syncthreads();
        for (int64_t b = threadIdx.x; b < nbins; b += 256)
            if (s_hist[b]) atomicAdd(&hist[b], (unsigned long long)s_hist[b]);
    }
    if (PASS < 2) {
        __shared__ double s_sum[256];
        __shared__ double s_sq[256];
        __shared__ float s_mn[256], s_mx[256];
        __shared__ long long s_cnt[256];
        __syncthreads();
        s_sum[threadIdx.x] = sum; s_sq[threadIdx.x] = sumsq; s_mn[threadIdx.x] = mn; s_mx[threadIdx.x] = mx; s_cnt[threadIdx.x] = cnt;
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
            CenStats o;
            o.sum = s_sum[0]; o.sumsq = s_sq[0]; o.mn = s_mn[0]; o.mx = s_mx[0]; o.cnt = s_cnt[0];
            partial[blockIdx.x] = o;
        }
    }
}

__global__ void __launch_bounds__(256) k_center_fold(const CenStats* __restrict__ partial, int64_t ntiles, CenStats* __restrict__ out) {
    double sum = 0.0, sq = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    for (int64_t t = threadIdx.x; t < ntiles; t += 256) {
        const CenStats p = partial[t];
        sum += p.sum; sq += p.sumsq; mn = fminf(mn, p.mn); mx = fmaxf(mx, p.mx); cnt += p.cnt;
    }
    __shared__ double s_sum[256];
    __shared__ double s_sq[256];
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
        CenStats o;
        o.sum = s_sum[0]; o.sumsq = s_sq[0]; o.mn = s_mn[0]; o.mx = s_mx[0]; o.cnt = s_cnt[0];
        *out = o;
    }
}

// Stage 1 of detect_center: h_out = {count_valid, r0, r1, min, max, mean, var} of the rank-trimmed samples.
// Leaves the tile prefix in the arena for stage 2 (urh_center_histogram must follow immediately).
extern "C" int urh_center_stats(urh_ctx* ctx, const float* d_x, int64_t n, int64_t max_size, double* h_out) {
    for (int i = 0; i < 7; i++) h_out[i] = 0.0;
    if (n <= 0) return URH_OK;
    urh_arena_reset(ctx);
    const int64_t ntiles = urh_div_up(n, CEN_TILE);
    int64_t* prefix;
    int64_t* d_total;
    CenStats* partial;
    CenStats* folded;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles + 1, &prefix));
    URH_CHECK(urh_arena(ctx, 4, &d_total));
    URH_CHECK(urh_arena(ctx, (size_t)ctx->sm_count * 8 + 8, &partial));
    URH_CHECK(urh_arena(ctx, 2, &folded));
    URH_LAUNCH(ctx, k_count_valid, (unsigned)ntiles, 256, 0, d_x, n, prefix);
    URH_CHECK((urhscan::device_scan<int64_t, urhscan::AddI64>(ctx, prefix, ntiles, urhscan::AddI64(), (int64_t)0, true, d_total)));
    URH_CUDA(ctx, cudaMemcpyAsync(prefix + ntiles, d_total, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    int64_t total = 0;
    URH_CHECK(urh_read_i64(ctx, d_total, 1, &total));
    // rect[int(0.05 * len(rect)) : int(0.95 * len(rect))]  (Python float arithmetic, truncation)
    int64_t r0 = (int64_t)(0.05 * (double)total), r1 = (int64_t)(0.95 * (double)total);
    if (max_size >= 0 && r1 - r0 > max_size) r1 = r0 + max_size;
    h_out[0] = (double)total; h_out[1] = (double)r0; h_out[2] = (double)r1;
    if (r1 <= r0) return URH_OK;
    const unsigned gs = (unsigned)min(ntiles, (int64_t)ctx->sm_count * 8);
    URH_LAUNCH(ctx, (k_center_pass<0>), gs, 256, 0, d_x, n, prefix, r0, r1, 0.0, 0.0, 1.0, (int64_t)0, partial, nullptr, ntiles, 0);
    URH_LAUNCH(ctx, k_center_fold, 1, 256, 0, partial, (int64_t)gs, folded);
    CenStats st;
    URH_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, folded, sizeof(CenStats), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(&st, ctx->h_mail, sizeof(st));
    const double mean = st.sum / (double)st.cnt;
    // population variance from the double sums (np.var semantics; the reference's float32 pairwise result differs ~1e-7)
    CenStats sv;
    sv.sum = st.sumsq - (double)st.cnt * mean * mean;
    if (sv.sum < 0.0) sv.sum = 0.0;
    h_out[3] = (double)st.mn; h_out[4] = (double)st.mx; h_out[5] = mean; h_out[6] = sv.sum / (double)st.cnt;
    return URH_OK;
}

// Stage 2: counts for edges hmin + k*hstep, k = 0..nbins (np.arange), over the same rank-trimmed samples.
extern "C" int urh_center_histogram(urh_ctx* ctx, const float* d_x, int64_t n, int64_t r0, int64_t r1, double hmin,
                                    double hstep, int64_t nbins, int64_t* h_hist) {
    if (nbins <= 0) return URH_OK;
    urh_arena_reset(ctx);
    const int64_t ntiles = urh_div_up(n, CEN_TILE);
    int64_t* prefix;
    int64_t* d_total;
    unsigned long long* hist;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles + 1, &prefix));
    URH_CHECK(urh_arena(ctx, 4, &d_total));
    URH_CHECK(urh_arena(ctx, (size_t)nbins, &hist));
    URH_LAUNCH(ctx, k_count_valid, (unsigned)ntiles, 256, 0, d_x, n, prefix);
    URH_CHECK((urhscan::device_scan<int64_t, urhscan::AddI64>(ctx, prefix, ntiles, urhscan::AddI64(), (int64_t)0, true, d_total)));
    URH_CUDA(ctx, cudaMemcpyAsync(prefix + ntiles, d_total, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    URH_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(unsigned long long), ctx->stream));
    const int in_smem = nbins <= 12000 ? 1 : 0;
    const unsigned gs = (unsigned)min(ntiles, (int64_t)ctx->sm_count * 8);
    URH_LAUNCH(ctx, (k_center_pass<2>), gs, 256, in_smem ? (size_t)nbins * sizeof(unsigned int) : 0, d_x, n, prefix, r0, r1, 0.0, hmin,
               hstep, nbins, nullptr, hist, ntiles, in_smem);
    URH_CUDA(ctx, cudaMemcpyAsync(h_hist, hist, (size_t)nbins * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return URH_OK;
}

// ---- run tables for the segmenter and the plateau RLE ----------------------------------------------------------------
// mode 0: class = x > thr (segment_messages, tolerance 9 <=> 10 consecutive samples, auto_interpretation.pyx:69)
// mode 1: class = x <= thr ? 0 : 1 (get_plateau_lengths, tolerance 0: every run start)
// Candidates (position, class) are returned to the host (they are few); *h_last = {last_cls, last_len, first_cls}.
template <typename SRC, typename T>
static int run_table(urh_ctx* ctx, const T* d_x, int64_t n, float thr, int tol, int64_t** h_pos, int16_t** h_cls,
                     int64_t* count, int64_t* h_last) {
    urh_arena_reset(ctx);
    UrhClassify cls;
    memset(&cls, 0, sizeof(cls));
    cls.noise_value = 0.0f;
    cls.order = 2;
    cls.thr[0] = thr;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int cap = URH_TILE / (tol + 1) + 2;
    UrhTileSummary* tiles;
    uint32_t* staging;
    URH_CHECK(urh_arena(ctx, (size_t)ntiles, &tiles));
    URH_CHECK(urh_arena(ctx, (size_t)ntiles * cap, &staging));
    const unsigned grid = (unsigned)urh_div_up(ntiles, URH_WARPS_PER_BLOCK);
    const int vec_in = (((uintptr_t)d_x % (2 * sizeof(T))) == 0) ? 1 : 0;
    URH_LAUNCH(ctx, (k_dense_f32<SRC, T>), grid, URH_WARPS_PER_BLOCK * 32, 0, d_x, n, vec_in, cls, tol, tiles, staging, cap,
               (int16_t*)nullptr, 0);
    UrhCandidates cand;
    URH_CHECK(urh_collect_candidates(ctx, n, tol, tiles, staging, cap, &cand));
    *count = cand.count;
    h_last[0] = cand.last_cls;
    h_last[1] = cand.last_len;
    *h_pos = nullptr;
    *h_cls = nullptr;
    if (cand.count > 0) {
        *h_pos = (int64_t*)malloc((size_t)cand.count * sizeof(int64_t));
        *h_cls = (int16_t*)malloc((size_t)cand.count * sizeof(int16_t));
        URH_CUDA(ctx, cudaMemcpyAsync(*h_pos, cand.pos, (size_t)cand.count * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CUDA(ctx, cudaMemcpyAsync(*h_cls, cand.cls, (size_t)cand.count * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return URH_OK;
}

// segment_messages_from_magnitudes (auto_interpretation.pyx:55-111).  d_mags float32 (is_f64=0) or float64.
// h_segments receives (start, end) pairs, capacity `cap` pairs; *k = number of messages (may exceed cap: call again).
extern "C" int urh_segment_messages(urh_ctx* ctx, const void* d_mags, int is_f64, int64_t n, float noise_threshold,
                                    int64_t* h_segments, int64_t cap, int64_t* k) {
    *k = 0;
    if (n <= 0) return URH_OK;
    int64_t* pos = nullptr;
    int16_t* cl = nullptr;
    int64_t count = 0, last[2];
    float first = 0.f;
    if (is_f64) {
        double f0;
        URH_CUDA(ctx, cudaMemcpyAsync(&f0, d_mags, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CHECK((run_table<SrcAbove, double>(ctx, (const double*)d_mags, n, noise_threshold, 9, &pos, &cl, &count, last)));
        first = (f0 > (double)noise_threshold) ? 1.f : 0.f;
    } else {
        float f0;
        URH_CUDA(ctx, cudaMemcpyAsync(&f0, d_mags, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        URH_CHECK((run_table<SrcAbove, float>(ctx, (const float*)d_mags, n, noise_threshold, 9, &pos, &cl, &count, last)));
        first = (f0 > noise_threshold) ? 1.f : 0.f;
    }
    // host tail: the two-state machine over the (few) runs of >= 10 samples
    int state = first > 0.f ? 1 : 0;
    int64_t start = 0, m = 0;
    for (int64_t j = 0; j < count; j++) {
        if (cl[j] == state) continue;
        const int64_t p = pos[j];  // index of the 10th consecutive sample of the opposite class
        if (state == 1) {
            if (m < cap) { h_segments[2 * m] = start; h_segments[2 * m + 1] = p - 10; }
            m++;
            state = 0;
        } else {
            start = p - 10;
            state = 1;
        }
    }
    if (state == 1) {
        const int64_t conseq_below = (last[0] == 0) ? last[1] : 0;
        if (start < n - conseq_below) {
            if (m < cap) { h_segments[2 * m] = start; h_segments[2 * m + 1] = n - conseq_below; }
            m++;
        }
    }
    free(pos);
    free(cl);
    *k = m;
    return URH_OK;
}

// get_plateau_lengths (auto_interpretation.pyx:179-208): h_out capacity `cap`; *k = number of plateaus.
extern "C" int urh_plateau_lengths(urh_ctx* ctx, const float* d_rect, int64_t n, float center, int percentage,
                                   uint64_t* h_out, int64_t cap, int64_t* k) {
    *k = 0;
    if (n <= 0) return URH_OK;
    int64_t* pos = nullptr;
    int16_t* cl = nullptr;
    int64_t count = 0, last[2];
    URH_CHECK((run_table<SrcCenter, float>(ctx, d_rect, n, center, 0, &pos, &cl, &count, last)));
    // candidates with tolerance 0 are the run starts (the first one is position 0)
    const uint64_t limit = (uint64_t)percentage * (uint64_t)n / 100;
    uint64_t sum = 0;
    int64_t m = 0;
    for (int64_t j = 1; j < count; j++) {
        // the reference checks `current_sum >= limit` at the top of every sample iteration, i.e. before a
        // boundary at pos[j] can append the run that ends there
        if (sum >= limit) break;
        const uint64_t len = (uint64_t)(pos[j] - pos[j - 1]);
        if (m < cap) h_out[m] = len;
        m++;
        sum += len;
    }
    if (limit == 0) m = 0;
    free(pos);
    free(cl);
    *k = m;
    return URH_OK;
}

// ---- median filter (auto_interpretation.pyx:211-240) ---------------------------------------------------------------------
// window [i, i+k) truncated at the end; values converted to float32 first; result = sorted[k'//2]
__global__ void k_median(const double* __restrict__ x, int64_t n, int k, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float buf[64];
    int kk = k;
    if (i + kk > n) kk = (int)(n - i);
    for (int j = 0; j < kk; j++) {
        const float v = (float)x[i + j];
        int p = j;
        while (p > 0 && buf[p - 1] > v) { buf[p] = buf[p - 1]; p--; }
        buf[p] = v;
    }
    out[i] = buf[kk / 2];
}

extern "C" int urh_median_filter(urh_ctx* ctx, const double* d_x, int64_t n, unsigned int k, float* d_out) {
    if (n <= 0) return URH_OK;
    if (k == 0 || k > 64) URH_FAIL(ctx, URH_ERR_INVALID, "median_filter: k must be in 1..64");
    URH_LAUNCH(ctx, k_median, (unsigned)urh_div_up(n, 128), 128, 0, d_x, n, (int)k, d_out);
    return URH_OK;
}

// ---- arr2decibel (util.pyx:38-48): 10.0f * log10f(re*re + im*im), float32 ------------------------------------------
__global__ void k_decibel(const float2* __restrict__ x, int64_t count, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float2 v = x[i];
        out[i] = __fmul_rn(10.0f, log10f(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))));
    }
}

extern "C" int urh_arr2decibel(urh_ctx* ctx, const float* d_complex, int64_t count, float* d_out) {
    if (count <= 0) return URH_OK;
    const unsigned grid = (unsigned)min((int64_t)ctx->sm_count * 16, urh_div_up(count, 256));
    URH_LAUNCH(ctx, k_decibel, grid, 256, 0, (const float2*)d_complex, count, d_out);
    return URH_OK;
}

// =====================================================================================================
// detect_center fed by the dense pass: the demodulator already produced per-tile {count, min, max, sum, sumsq} of the
// samples detect_center keeps (UrhTileStats), so the trimmed statistics need no pass over qad at all (only the two
// tiles that contain the 5 % / 95 % rank cuts are re-read) and the histogram is the single extra pass.
// =====================================================================================================
__global__ void k_tile_counts(const UrhTileStats* __restrict__ ts, int64_t ntiles, int64_t* __restrict__ prefix) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntiles) prefix[t] = ts[t].cnt;
}

// rank-exact partial statistics of one tile (used for the <= 2 tiles cut by the rank window); one block per listed tile
__global__ void __launch_bounds__(256) k_edge_tile_stats(const float* __restrict__ x, int64_t n, const int64_t* __restrict__ prefix,
                                                        const int64_t* __restrict__ edge_tiles, int64_t r0, int64_t r1,
                                                        CenStats* __restrict__ out) {
    const int64_t t = edge_tiles[blockIdx.x];
    CenStats o;
    o.sum = 0.0; o.sumsq = 0.0; o.mn = INFINITY; o.mx = -INFINITY; o.cnt = 0;
    if (t >= 0) {
        const int per = URH_TILE / 256;
        const int64_t base = t * URH_TILE + (int64_t)threadIdx.x * per;
        float v[per];
        int mine = 0;
#pragma unroll
        for (int j = 0; j < per; j++) {
            v[j] = (base + j < n) ? x[base + j] : -5.0f;
            mine += (v[j] > -4.0f) ? 1 : 0;
        }
        __shared__ int s_pre[256];
        s_pre[threadIdx.x] = mine;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            int add = 0;
            if (threadIdx.x >= off) add = s_pre[threadIdx.x - off];
            __syncthreads();
            s_pre[threadIdx.x] += add;
            __syncthreads();
        }
        int64_t rank = prefix[t] + s_pre[threadIdx.x] - mine;
#pragma unroll
        for (int j = 0; j < per; j++) {
            if (v[j] > -4.0f) {
                if (rank >= r0 && rank < r1) {
                    o.sum += (double)v[j];
                    o.sumsq += (double)v[j] * (double)v[j];
                    o.mn = fminf(o.mn, v[j]);
                    o.mx = fmaxf(o.mx, v[j]);
                    o.cnt++;
                }
                rank++;
            }
        }
    }
    __shared__ double s_sum[256], s_sq[256];
    __shared__ float s_mn[256], s_mx[256];
    __shared__ long long s_cnt[256];
    s_sum[threadIdx.x] = o.sum; s_sq[threadIdx.x] = o.sumsq; s_mn[threadIdx.x] = o.mn; s_mx[threadIdx.x] = o.mx; s_cnt[threadIdx.x] = o.cnt;
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
        out[blockIdx.x] = r;
    }
}

// interior tiles (entirely inside the rank window): fold the dense pass's partials; grid-stride, one partial per block
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
        partial[blockIdx.x] = r;
    }
}

// histogram over the dense pass's tiles: interior tiles need no rank bookkeeping.  A demodulated capture piles its
// samples onto a handful of bins, so lanes that hit the same bin are merged (__match_any_sync) into one shared-memory atomic.
__global__ void __launch_bounds__(256) k_hist_tiles(const float* __restrict__ x, int64_t n, const int64_t* __restrict__ prefix,
                                                   int64_t ntiles, int64_t r0, int64_t r1, double hmin, double hstep, int64_t nbins,
                                                   unsigned long long* __restrict__ hist, int hist_in_smem) {
    extern __shared__ unsigned int s_hist[];
    __shared__ int s_pre[256];
    const int lane = threadIdx.x & 31;
    if (hist_in_smem) {
        for (int64_t b = threadIdx.x; b < nbins; b += 256) s_hist[b] = 0u;
        __syncthreads();
    }
    const double last_edge = hmin + (double)nbins * hstep;
    const double inv = 1.0 / hstep;
    // bin of one value, -1 when it does not count (np.histogram: half-open bins, the last one closed)
    auto bin_of = [&](float f, bool counts) -> int {
        if (!counts) return -1;
        const double a = (double)f;
        if (!(a >= hmin && a <= last_edge)) return -1;
        int64_t k = (int64_t)floor((a - hmin) * inv);
        if (k < 0) k = 0;
        if (k > nbins - 1) k = nbins - 1;
        while (k > 0 && a < hmin + (double)k * hstep) k--;
        while (k < nbins - 1 && a >= hmin + (double)(k + 1) * hstep) k++;
        return (int)k;
    };
    // called by all 32 lanes of a warp together
    auto put = [&](int k) {
        const unsigned peers = __match_any_sync(URH_FULL_MASK, k);
        if (k >= 0 && lane == __ffs(peers) - 1) {
            if (hist_in_smem) atomicAdd(&s_hist[k], (unsigned)__popc(peers));
            else atomicAdd(&hist[k], (unsigned long long)__popc(peers));
        }
    };
    const int per = URH_TILE / 256;
    const bool vec = (((uintptr_t)x) & 15) == 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t a = prefix[t], b = prefix[t + 1];
        if (b <= a || b <= r0 || a >= r1) continue;  // block-uniform
        const int64_t base = t * URH_TILE;
        if (a >= r0 && b <= r1) {
            // interior: coalesced, every kept sample counts
            if (vec && base + URH_TILE <= n) {
                const float4* p4 = (const float4*)(x + base);
                float4 v[URH_TILE / 1024];
#pragma unroll
                for (int j = 0; j < URH_TILE / 1024; j++) v[j] = __ldg(p4 + j * 256 + threadIdx.x);
#pragma unroll
                for (int j = 0; j < URH_TILE / 1024; j++) {
                    put(bin_of(v[j].x, v[j].x > -4.0f));
                    put(bin_of(v[j].y, v[j].y > -4.0f));
                    put(bin_of(v[j].z, v[j].z > -4.0f));
                    put(bin_of(v[j].w, v[j].w > -4.0f));
                }
            } else {
                for (int j = threadIdx.x; j < URH_TILE; j += 256) {
                    const int64_t i = base + j;
                    const float f = (i < n) ? x[i] : -5.0f;
                    put(bin_of(f, f > -4.0f));
                }
            }
        } else {
            float v[per];
            int mine = 0;
#pragma unroll
            for (int j = 0; j < per; j++) {
                const int64_t i = base + (int64_t)threadIdx.x * per + j;
                v[j] = (i < n) ? x[i] : -5.0f;
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
            int64_t rank = a + s_pre[threadIdx.x] - mine;
#pragma unroll
            for (int j = 0; j < per; j++) {
                const bool kept = v[j] > -4.0f;
                put(bin_of(v[j], kept && rank >= r0 && rank < r1));
                rank += kept ? 1 : 0;
            }
            __syncthreads();
        }
    }
    if (hist_in_smem) {
        __syncthreads();
        for (int64_t b = threadIdx.x; b < nbins; b += 256)
            if (s_hist[b]) atomicAdd(&hist[b], (unsigned long long)s_hist[b]);
    }
}

// the (at most two) tiles the rank window [r0, r1) cuts: tile t holds ranks [prefix[t], prefix[t+1]).  edges[q] = tile index
// or -1 (no such tile / the window covers it completely, so the interior kernel takes it / same tile as edges[0]).
__global__ void k_find_edge_tiles(const int64_t* __restrict__ prefix, int64_t ntiles, int64_t r0, int64_t r1, int64_t* __restrict__ edges) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int64_t a = prefix[t], b = prefix[t + 1];
    if (b <= a) return;
    const bool covered = a >= r0 && b <= r1;
    if (a <= r0 && r0 < b) edges[0] = covered ? -1 : t;
    if (a <= r1 - 1 && r1 - 1 < b && !(a <= r0 && r0 < b)) edges[1] = covered ? -1 : t;
}

// Rank prefix over the tile table the demodulator produced; leaves {ts, prefix, n} in ctx for the window / histogram calls.
int urh_center_tiles_begin(urh_ctx* ctx, const UrhTileStats* ts, int64_t n, int64_t* h_total) {
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
    return URH_OK;
}

// {count, min, max, sum, sumsq} of the kept samples whose LOCAL rank is in [r0, r1): interior tiles from the table, the
// two cut tiles re-read from qad.  A shard passes the global window minus its rank offset (clamped to its own count).
extern "C" int urh_center_window_stats(urh_ctx* ctx, const float* d_qad, int64_t n, int64_t r0, int64_t r1, double* h_out5) {
    h_out5[0] = 0.0; h_out5[1] = INFINITY; h_out5[2] = -INFINITY; h_out5[3] = 0.0; h_out5[4] = 0.0;
    if (!ctx->center_prefix || ctx->center_n != n) URH_FAIL(ctx, URH_ERR_INVALID, "urh_afp_demod_tiles must precede urh_center_window_stats");
    if (r1 <= r0) return URH_OK;
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    const int64_t* prefix = (const int64_t*)ctx->center_prefix;
    const UrhTileStats* ts = (const UrhTileStats*)ctx->center_ts;
    int64_t* d_edges;
    CenStats* partial;
    CenStats* folded;
    const int nb = ctx->sm_count * 2;
    URH_CHECK(urh_arena(ctx, 4, &d_edges));
    URH_CHECK(urh_arena(ctx, (size_t)nb + 4, &partial));
    URH_CHECK(urh_arena(ctx, 2, &folded));
    const int64_t none[2] = {-1, -1};
    URH_CUDA(ctx, cudaMemcpyAsync(d_edges, none, sizeof(none), cudaMemcpyHostToDevice, ctx->stream));
    URH_LAUNCH(ctx, k_find_edge_tiles, (unsigned)urh_div_up(ntiles, 256), 256, 0, prefix, ntiles, r0, r1, d_edges);
    URH_LAUNCH(ctx, k_interior_tile_stats, nb, 256, 0, ts, prefix, ntiles, r0, r1, partial);
    URH_LAUNCH(ctx, k_edge_tile_stats, 2, 256, 0, d_qad, n, prefix, d_edges, r0, r1, partial + nb);
    URH_LAUNCH(ctx, k_center_fold, 1, 256, 0, partial, (int64_t)nb + 2, folded);
    CenStats st;
    URH_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, folded, sizeof(CenStats), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(&st, ctx->h_mail, sizeof(st));
    h_out5[0] = (double)st.cnt; h_out5[1] = (double)st.mn; h_out5[2] = (double)st.mx; h_out5[3] = st.sum; h_out5[4] = st.sumsq;
    return URH_OK;
}

// histogram pass that goes with urh_afp_demod_stats (uses the tile prefix it left in the arena)
extern "C" int urh_center_histogram_tiles(urh_ctx* ctx, const float* d_qad, int64_t n, int64_t r0, int64_t r1, double hmin,
                                          double hstep, int64_t nbins, int64_t* h_hist) {
    if (nbins <= 0) return URH_OK;
    if (!ctx->center_prefix || ctx->center_n != n) URH_FAIL(ctx, URH_ERR_INVALID, "urh_afp_demod_tiles must precede urh_center_histogram_tiles");
    const int64_t ntiles = urh_div_up(n, URH_TILE);
    unsigned long long* hist;
    URH_CHECK(urh_arena(ctx, (size_t)nbins, &hist));
    URH_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(unsigned long long), ctx->stream));
    const int in_smem = nbins <= 12000 ? 1 : 0;
    const unsigned gs = (unsigned)min(ntiles, (int64_t)ctx->sm_count * 8);
    URH_LAUNCH(ctx, k_hist_tiles, gs, 256, in_smem ? (size_t)nbins * sizeof(unsigned int) : 0, d_qad, n, (const int64_t*)ctx->center_prefix,
               ntiles, r0, r1, hmin, hstep, nbins, hist, in_smem);
    URH_CUDA(ctx, cudaMemcpyAsync(h_hist, hist, (size_t)nbins * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    URH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return URH_OK;
}
